# 2-GPU: dist tests + bench N=2 (200-iteration steps, shortened all-reduce path)
set -x
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
timeout 900 python -m pytest tests/test_gpu_dist.py -x -q 2>&1 | tail -8 > gpurun_out/r2_c11_pytest_dist.log
cat gpurun_out/r2_c11_pytest_dist.log
timeout 600 $TR --master-port 29651 bench.py --gpus 2 --steps 5 --warmup 3 > gpurun_out/r2_c11_bench_n2.json 2> gpurun_out/r2_c11_bench_n2.err
tail -3 gpurun_out/r2_c11_bench_n2.err
python - <<'PY'
import json
d=json.loads(open("gpurun_out/r2_c11_bench_n2.json").read().strip().splitlines()[-1])
c5=d.get("cfg5",{})
print("%.1f it/s e2e %.1f launches %d"%(d["value"],d["e2e"]["value"],d["gpu_launches"]), d["roofline"].get("kernels"), "cfg5 %.1f"%c5.get("value",0), c5.get("kernels"), d.get("parity"), c5.get("parity"))
PY
