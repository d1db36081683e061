# 1-GPU: full suite, bench (zero-copy report, batch 32), block bench + launch list, solver bench incl. siblings
set -x
mkdir -p gpurun_out
timeout 1800 python -m pytest tests -m gpu -q --deselect tests/test_gpu_dist.py 2>&1 | tail -15 > gpurun_out/r2_c8_pytest.log
cat gpurun_out/r2_c8_pytest.log
timeout 600 python bench.py --steps 5 --warmup 3 > gpurun_out/r2_c8_bench.json 2> gpurun_out/r2_c8_bench.err; tail -3 gpurun_out/r2_c8_bench.err
timeout 600 python profiles/bench_block.py 8 16 32 > gpurun_out/r2_c8_block.jsonl 2> gpurun_out/r2_c8_block.err; cat gpurun_out/r2_c8_block.jsonl
timeout 600 python profiles/bench_solvers.py gmres bicgstab minres siblings > gpurun_out/r2_c8_solvers.jsonl 2> gpurun_out/r2_c8_solvers.err; cat gpurun_out/r2_c8_solvers.jsonl
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"spmm|panel|relayout|ew_kernel|dot_kernel" -s 200 -c 150 --csv --log-file gpurun_out/r2_launches_block_p32.csv python profiles/bench_block.py 32 > gpurun_out/r2_c8_launches_block.log 2>&1
python - <<'PY'
import json
d=json.loads(open("gpurun_out/r2_c8_bench.json").read().strip().splitlines()[-1])
print("value %.1f frac %.4f e2e %.1f launches %d"%(d["value"],d["roofline"]["frac"],d["e2e"]["value"],d["gpu_launches"]), d["roofline"]["kernels"], d.get("parity"), [ (e.get("solver"), round(e.get("value",0),1), round(e.get("roofline",{}).get("frac",0),3)) for e in d.get("extra",[])], d.get("cfg5",{}).get("value"))
PY
