# 2-GPU run: dist tests, bench N=2 (default build), then plan / gather-depth variants of the persistent dist kernel
set -x
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
timeout 900 python -m pytest tests/test_gpu_dist.py -x -q 2>&1 | tail -25 > gpurun_out/r2_c3_pytest_dist.log
cat gpurun_out/r2_c3_pytest_dist.log
timeout 600 $TR --master-port 29611 bench.py --gpus 2 --steps 5 --warmup 3 > gpurun_out/r2_c3_bench_n2.json 2> gpurun_out/r2_c3_bench_n2.err
tail -3 gpurun_out/r2_c3_bench_n2.err
i=0
for cfg in "KB200_GATHER_DEPTH=4" "KB200_CTAS_PER_SM=2 KB200_STAGES=3" "KB200_CTAS_PER_SM=2 KB200_STAGES=4" "KB200_PERSIST=0"; do
  i=$((i+1))
  env $cfg timeout 600 $TR --master-port $((29620+i)) bench.py --gpus 2 --steps 5 --warmup 3 --no-cpu > gpurun_out/r2_c3_bench_n2_v$i.json 2> gpurun_out/r2_c3_bench_n2_v$i.err
  echo "$cfg" > gpurun_out/r2_c3_bench_n2_v$i.cfg
done
python - <<'PY'
import json,glob
for f in sorted(glob.glob("gpurun_out/r2_c3_bench_n2*.json")):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
        c5=d.get("cfg5",{})
        print(f, "%.1f it/s"%d["value"], d["roofline"].get("kernels"), "cfg5 %.1f"%c5.get("value",0), c5.get("kernels"), (d.get("parity") or {}).get("ok"))
    except Exception as e:
        print(f, "ERR", e)
PY
