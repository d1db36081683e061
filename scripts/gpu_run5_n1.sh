# 1-GPU: full single-GPU test suite, ncu captures (block panel kernels p = 16 / 32, cg_persist), launch list
set -x
mkdir -p gpurun_out
timeout 1800 python -m pytest tests -m gpu -q --deselect tests/test_gpu_dist.py 2>&1 | tail -15 > gpurun_out/r2_c5_pytest.log
cat gpurun_out/r2_c5_pytest.log
timeout 600 python profiles/bench_block.py 8 16 32 > gpurun_out/r2_c5_block.jsonl 2> gpurun_out/r2_c5_block.err; cat gpurun_out/r2_c5_block.jsonl
for p in 16 32; do
  timeout 900 ncu --set full --clock-control none --import-source on -k regex:panel_mma_kernel -s 20 -c 3 -f -o gpurun_out/r2_ncu_block_p$p python profiles/bench_block.py $p > gpurun_out/r2_c5_ncu_block_p$p.log 2>&1
done
timeout 900 ncu --set full --clock-control none --import-source on -k regex:cg_persist -s 2 -c 1 -f -o gpurun_out/r2_ncu_cg_persist python bench.py --steps 1 --warmup 1 --no-cpu --no-extra --no-cfg5 > gpurun_out/r2_c5_ncu_cg.log 2>&1
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/r2_launches_cg.csv python bench.py --steps 2 --warmup 1 --no-cpu --no-extra --no-cfg5 > gpurun_out/r2_c5_launches.log 2>&1
ls -la gpurun_out/*.ncu-rep
