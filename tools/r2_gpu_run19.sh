#!/bin/bash
# round-2 GPU session 19: mask image builder with one thread per (unit, 64-k block): smoke, every GPU test file except the long
# parity file (its cassava-scale check is repeated inside the bench run), bench
mkdir -p gpurun_out
python __graft_entry__.py --smoke > gpurun_out/r2u_smoke.txt 2>&1; echo "smoke rc=$?"; tail -1 gpurun_out/r2u_smoke.txt
timeout 300 python -m pytest tests/test_gpu_accuracy.py tests/test_gpu_features.py tests/test_gpu_fuzz.py tests/test_gpu_rshim.py tests/test_gpu_host_api.py -m gpu -x -q > gpurun_out/r2u_pytest.txt 2>&1; echo "pytest rc=$?"; tail -2 gpurun_out/r2u_pytest.txt
timeout 400 python bench.py --steps 20 --warmup 5 > gpurun_out/r2u_bench.json 2> gpurun_out/r2u_bench.err; echo "bench rc=$?"; head -c 300 gpurun_out/r2u_bench.json; echo
