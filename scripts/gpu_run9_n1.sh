set -x
mkdir -p gpurun_out
timeout 1800 python -m pytest tests -m gpu -x -q --tb=short --deselect tests/test_gpu_dist.py > gpurun_out/r2_c9_pytest.log 2>&1
tail -40 gpurun_out/r2_c9_pytest.log
