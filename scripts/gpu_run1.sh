set -x
mkdir -p gpurun_out
nvidia-smi -L
# 1. the CG paths first (persistent kernel is new)
timeout 600 python -m pytest tests/test_gpu_solvers.py -x -q 2>&1 | tail -15 > gpurun_out/r2_c1_pytest_solvers.log
cat gpurun_out/r2_c1_pytest_solvers.log
# 2. bench: persistent vs two-launch
timeout 600 python bench.py --steps 5 --warmup 3 > gpurun_out/r2_c1_bench.json 2> gpurun_out/r2_c1_bench.err; tail -3 gpurun_out/r2_c1_bench.err
KB200_PERSIST=0 timeout 300 python bench.py --steps 5 --warmup 3 --no-cpu --no-extra --no-cfg5 > gpurun_out/r2_c1_bench_2launch.json 2> gpurun_out/r2_c1_bench_2launch.err
# 3. other solvers (clamped gathers + min-blocks hint)
timeout 600 python profiles/bench_solvers.py gmres bicgstab minres > gpurun_out/r2_c1_solvers.jsonl 2> gpurun_out/r2_c1_solvers.err
# 4. rest of GPU suite
timeout 1500 python -m pytest tests -m gpu -x -q --deselect tests/test_gpu_solvers.py 2>&1 | tail -15 > gpurun_out/r2_c1_pytest_rest.log
cat gpurun_out/r2_c1_pytest_rest.log
