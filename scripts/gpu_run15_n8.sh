# 8-GPU: one bench line at N = 8 with the final code
set -x
mkdir -p gpurun_out
timeout 600 python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1 --nproc-per-node 8 --master-port 29681 bench.py --gpus 8 --steps 5 --warmup 3 > gpurun_out/r2_c15_bench_n8.json 2> gpurun_out/r2_c15_bench_n8.err
tail -3 gpurun_out/r2_c15_bench_n8.err
python - <<'PY'
import json
d=json.loads(open("gpurun_out/r2_c15_bench_n8.json").read().strip().splitlines()[-1])
c5=d.get("cfg5",{})
print("%.1f it/s e2e %.1f launches %d"%(d["value"],d["e2e"]["value"],d["gpu_launches"]), d["roofline"].get("kernels"), "cfg5 %.1f"%c5.get("value",0), c5.get("kernels"), d.get("parity"), c5.get("parity"))
PY
