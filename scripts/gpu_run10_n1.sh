# 1-GPU final: smoke(), full suite, default bench (the driver's command), --impl reference arm, BiCGSTAB f32 ncu counters,
# launch list of the default bench
set -x
mkdir -p gpurun_out
timeout 600 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > gpurun_out/r2_c10_smoke.log 2>&1; tail -2 gpurun_out/r2_c10_smoke.log
timeout 1800 python -m pytest tests -m gpu -q --tb=short > gpurun_out/r2_c10_pytest.log 2>&1; tail -4 gpurun_out/r2_c10_pytest.log
timeout 900 python bench.py > gpurun_out/r2_c10_bench_default.json 2> gpurun_out/r2_c10_bench_default.err; tail -2 gpurun_out/r2_c10_bench_default.err
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/r2_c10_bench_reference.json 2> gpurun_out/r2_c10_bench_reference.err
timeout 900 ncu --set full --clock-control none -k regex:spmv_epi_tma -s 6 -c 2 -f -o gpurun_out/r2_ncu_bicgstab_spmv python profiles/bench_solvers.py bicgstab > gpurun_out/r2_c10_ncu_bicg.log 2>&1
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 40 --csv --log-file gpurun_out/r2_launches_bench.csv python bench.py --steps 2 --warmup 1 --no-cpu --no-extra --no-cfg5 > gpurun_out/r2_c10_launches.log 2>&1
for bt in 8 16 32; do KB200_BATCH=$bt timeout 300 python bench.py --steps 5 --warmup 3 --no-cpu --no-extra --no-cfg5 > gpurun_out/r2_c10_bench_batch$bt.json 2>/dev/null; done
timeout 300 python scripts/probe_solve_overhead.py > gpurun_out/r2_c10_probe.txt 2>&1; cat gpurun_out/r2_c10_probe.txt
python - <<'PY'
import json
for bt in (8,16,32):
    try:
        d=json.loads(open(f"gpurun_out/r2_c10_bench_batch{bt}.json").read().strip().splitlines()[-1]); k=d['roofline']['kernels']
        print('batch',bt,'%.1f it/s'%d['value'],'ms/step %.3f'%d['ms_per_step'],'phases %.1f %.1f us'%(1e3*k['phase_a']['ms'],1e3*k['phase_b']['ms']),'launches',d['gpu_launches'])
    except Exception as e: print(bt,'ERR',e)
d=json.loads(open("gpurun_out/r2_c10_bench_default.json").read().strip().splitlines()[-1])
print("value %.1f frac %.4f e2e %.1f launches %d"%(d["value"],d["roofline"]["frac"],d["e2e"]["value"],d["gpu_launches"]), d["roofline"]["kernels"], d.get("parity"), [ (e.get("solver"), round(e.get("value",0),1), round(e.get("roofline",{}).get("frac",0),3)) for e in d.get("extra",[])], d.get("cfg5",{}).get("value"), d.get("cpu_baseline"))
PY
