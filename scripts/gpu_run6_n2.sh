# 2-GPU: dist tests + bench N=2 after moving the halo staging to the consumer warps
set -x
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
timeout 900 python -m pytest tests/test_gpu_dist.py -x -q 2>&1 | tail -25 > gpurun_out/r2_c6_pytest_dist.log
cat gpurun_out/r2_c6_pytest_dist.log
timeout 600 $TR --master-port 29631 bench.py --gpus 2 --steps 5 --warmup 3 > gpurun_out/r2_c6_bench_n2.json 2> gpurun_out/r2_c6_bench_n2.err
tail -3 gpurun_out/r2_c6_bench_n2.err
KB200_GATHER_DEPTH=4 timeout 600 $TR --master-port 29632 bench.py --gpus 2 --steps 5 --warmup 3 --no-cpu > gpurun_out/r2_c6_bench_n2_d4.json 2> gpurun_out/r2_c6_bench_n2_d4.err
python - <<'PY'
import json,glob
for f in sorted(glob.glob("gpurun_out/r2_c6_bench_n2*.json")):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
        c5=d.get("cfg5",{})
        print(f, "%.1f it/s"%d["value"], d["roofline"].get("kernels"), "cfg5 %.1f"%c5.get("value",0), c5.get("kernels"), (d.get("parity") or {}).get("ok"))
    except Exception as e:
        print(f, "ERR", e)
PY
