# 2-GPU run: row-partitioned tests + bench (poisson215 with parity, cfg5 riding along)
set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_dist.py -x -q 2>&1 | tail -25 > gpurun_out/r2_c2_pytest_dist.log
cat gpurun_out/r2_c2_pytest_dist.log
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29601 bench.py --gpus 2 --steps 5 --warmup 3 > gpurun_out/r2_c2_bench_n2.json 2> gpurun_out/r2_c2_bench_n2.err
tail -5 gpurun_out/r2_c2_bench_n2.err; cat gpurun_out/r2_c2_bench_n2.json
KB200_PERSIST=0 timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29602 bench.py --gpus 2 --steps 5 --warmup 3 --no-cpu > gpurun_out/r2_c2_bench_n2_2launch.json 2> gpurun_out/r2_c2_bench_n2_2launch.err
cat gpurun_out/r2_c2_bench_n2_2launch.json
