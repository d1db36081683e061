# 8-GPU: bench at N = 8 (default + two-launch A/B) and N = 4; cfg5 and parity ride along
set -x
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
timeout 600 $TR --nproc-per-node 8 --master-port 29641 bench.py --gpus 8 --steps 5 --warmup 3 > gpurun_out/r2_c7_bench_n8.json 2> gpurun_out/r2_c7_bench_n8.err
tail -3 gpurun_out/r2_c7_bench_n8.err
timeout 600 $TR --nproc-per-node 4 --master-port 29642 bench.py --gpus 4 --steps 5 --warmup 3 > gpurun_out/r2_c7_bench_n4.json 2> gpurun_out/r2_c7_bench_n4.err
KB200_PERSIST=0 timeout 600 $TR --nproc-per-node 8 --master-port 29643 bench.py --gpus 8 --steps 5 --warmup 3 --no-cpu > gpurun_out/r2_c7_bench_n8_2launch.json 2> gpurun_out/r2_c7_bench_n8_2launch.err
python - <<'PY'
import json,glob
for f in sorted(glob.glob("gpurun_out/r2_c7_bench_n*.json")):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
        c5=d.get("cfg5",{})
        print(f, "%.1f it/s"%d["value"], d["roofline"].get("kernels"), "cfg5 %.1f"%c5.get("value",0), c5.get("kernels"), (d.get("parity") or {}).get("ok"), (c5.get("parity") or {}).get("ok"))
    except Exception as e:
        print(f, "ERR", e)
PY
