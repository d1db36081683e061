# 4-GPU: bench at N = 4 with the final code (interior ranks with two neighbours), parity blocks included
set -x
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
timeout 600 $TR --nproc-per-node 4 --master-port 29671 bench.py --gpus 4 --steps 5 --warmup 3 > gpurun_out/r2_c14_bench_n4.json 2> gpurun_out/r2_c14_bench_n4.err
tail -3 gpurun_out/r2_c14_bench_n4.err
python - <<'PY'
import json
d=json.loads(open("gpurun_out/r2_c14_bench_n4.json").read().strip().splitlines()[-1])
c5=d.get("cfg5",{})
print("%.1f it/s e2e %.1f launches %d"%(d["value"],d["e2e"]["value"],d["gpu_launches"]), d["roofline"].get("kernels"), "cfg5 %.1f"%c5.get("value",0), c5.get("kernels"), d.get("parity"), c5.get("parity"))
PY
