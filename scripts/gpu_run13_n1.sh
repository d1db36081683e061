# 1-GPU final validation: smoke, full GPU suite, the driver's default bench command and the reference arm
set -x
mkdir -p gpurun_out
timeout 600 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > gpurun_out/r2_c13_smoke.log 2>&1; tail -2 gpurun_out/r2_c13_smoke.log
timeout 1800 python -m pytest tests -m gpu -q --tb=short > gpurun_out/r2_c13_pytest.log 2>&1; tail -6 gpurun_out/r2_c13_pytest.log
timeout 900 python bench.py > gpurun_out/r2_c13_bench_default.json 2> gpurun_out/r2_c13_bench_default.err; tail -2 gpurun_out/r2_c13_bench_default.err
python - <<'PY'
import json
d=json.loads(open("gpurun_out/r2_c13_bench_default.json").read().strip().splitlines()[-1])
print("value %.1f frac %.4f e2e %.1f launches %d"%(d["value"],d["roofline"]["frac"],d["e2e"]["value"],d["gpu_launches"]), d["roofline"]["kernels"], d.get("parity"), [ (e.get("solver"), round(e.get("value",0),1), round(e.get("roofline",{}).get("frac",0),3)) for e in d.get("extra",[])], d.get("cfg5",{}).get("value"), d.get("cfg5",{}).get("parity"), d.get("clocks"))
PY
