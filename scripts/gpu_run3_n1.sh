# 1-GPU run: block GMRES tensor-core path (tests + bench), CG plan / gather-depth sweep, the remaining GPU tests
set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_block.py -x -q 2>&1 | tail -15 > gpurun_out/r2_c4_pytest_block.log
cat gpurun_out/r2_c4_pytest_block.log
timeout 600 python profiles/bench_block.py 8 16 32 > gpurun_out/r2_c4_block_mma.jsonl 2> gpurun_out/r2_c4_block_mma.err
KB200_BLOCK_MMA=0 timeout 600 python profiles/bench_block.py 8 16 32 > gpurun_out/r2_c4_block_simt.jsonl 2> gpurun_out/r2_c4_block_simt.err
KB200_BLOCK_MMA=8 timeout 300 python profiles/bench_block.py 8 > gpurun_out/r2_c4_block_mma8.jsonl 2> gpurun_out/r2_c4_block_mma8.err
cat gpurun_out/r2_c4_block_mma.jsonl gpurun_out/r2_c4_block_simt.jsonl gpurun_out/r2_c4_block_mma8.jsonl
i=0
for cfg in "KB200_GATHER_DEPTH=8" "KB200_GATHER_DEPTH=4" "KB200_CTAS_PER_SM=2 KB200_STAGES=3" "KB200_CTAS_PER_SM=2 KB200_STAGES=4" "KB200_CTAS_PER_SM=2 KB200_STAGES=4 KB200_GATHER_DEPTH=4"; do
  i=$((i+1))
  env $cfg timeout 300 python bench.py --steps 5 --warmup 3 --no-cpu --no-extra --no-cfg5 > gpurun_out/r2_c4_bench_v$i.json 2> gpurun_out/r2_c4_bench_v$i.err
  echo "$cfg" > gpurun_out/r2_c4_bench_v$i.cfg
done
python - <<'PY'
import json,glob
for f in sorted(glob.glob("gpurun_out/r2_c4_bench_v*.json")):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
        k=d["roofline"]["kernels"]
        print(f, open(f.replace(".json",".cfg")).read().strip(), "%.1f it/s frac %.4f"%(d["value"],d["roofline"]["frac"]), "A %.1f us B %.1f us"%(1e3*k["phase_a"]["ms"],1e3*k["phase_b"]["ms"]))
    except Exception as e:
        print(f, "ERR", e)
PY
timeout 1500 python -m pytest tests -m gpu -x -q --deselect tests/test_gpu_block.py --deselect tests/test_gpu_dist.py 2>&1 | tail -15 > gpurun_out/r2_c4_pytest_rest.log
cat gpurun_out/r2_c4_pytest_rest.log
