#!/bin/bash
# round-2 GPU session 17 (2 GPUs): the driver's own N = 2 launch of bench.py on the final tree (exercises the ranks-agree warm-up loop
# under NCCL), the 2-rank NCCL tests, and the 1x2 chromosome grid
mkdir -p gpurun_out
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29531 bench.py --gpus 2 --steps 10 --warmup 3 > gpurun_out/r2s_bench_2gpu.json 2> gpurun_out/r2s_bench_2gpu.err; echo "bench2 rc=$?"; head -c 300 gpurun_out/r2s_bench_2gpu.json; echo
timeout 300 python -m pytest tests/test_gpu_multirank.py -m gpu -x -q > gpurun_out/r2s_pytest_2gpu.txt 2>&1; echo "pytest rc=$?"; tail -2 gpurun_out/r2s_pytest_2gpu.txt
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29532 bench.py --gpus 2 --grid 1x2 --steps 10 --warmup 3 --modes auto --no-cached > gpurun_out/r2s_bench_2gpu_1x2.json 2> gpurun_out/r2s_bench_2gpu_1x2.err; echo "bench 1x2 rc=$?"; head -c 300 gpurun_out/r2s_bench_2gpu_1x2.json; echo
