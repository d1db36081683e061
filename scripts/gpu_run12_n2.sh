# 2-GPU box: CG-related single-GPU tests, dist tests, bench at N=1 (short) and N=2 after the barrier changes
set -x
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
timeout 900 python -m pytest tests/test_gpu_solvers.py tests/test_gpu_fused_phases.py tests/test_gpu_blockjacobi.py tests/test_gpu_dist.py tests/test_gpu_capi.py -x -q 2>&1 | tail -6 > gpurun_out/r2_c12_pytest.log
cat gpurun_out/r2_c12_pytest.log
timeout 600 python bench.py --steps 5 --warmup 3 --no-extra --no-cfg5 > gpurun_out/r2_c12_bench_n1.json 2> gpurun_out/r2_c12_bench_n1.err
timeout 600 $TR --master-port 29661 bench.py --gpus 2 --steps 5 --warmup 3 > gpurun_out/r2_c12_bench_n2.json 2> gpurun_out/r2_c12_bench_n2.err
tail -3 gpurun_out/r2_c12_bench_n2.err
python - <<'PY'
import json
for f in ("gpurun_out/r2_c12_bench_n1.json","gpurun_out/r2_c12_bench_n2.json"):
    d=json.loads(open(f).read().strip().splitlines()[-1])
    c5=d.get("cfg5",{})
    print(f, "%.1f it/s frac %.4f e2e %.1f"%(d["value"],d["roofline"]["frac"],d["e2e"]["value"]), d["roofline"].get("kernels"), "cfg5 %.1f"%c5.get("value",0), (d.get("parity") or {}).get("ok"), (d.get("parity") or {}).get("max_rel_dev"), (c5.get("parity") or {}).get("ok"))
PY
