#!/bin/bash
# round-2 GPU session 18: cache off = every operand image (the per-cross mask image too) rebuilt in every compute; bench line with
# the sustained-peak fraction; feature / accuracy / shim tests
mkdir -p gpurun_out
python __graft_entry__.py --smoke > gpurun_out/r2t_smoke.txt 2>&1; echo "smoke rc=$?"; tail -1 gpurun_out/r2t_smoke.txt
timeout 600 python bench.py --steps 20 --warmup 5 > gpurun_out/r2t_bench.json 2> gpurun_out/r2t_bench.err; echo "bench rc=$?"; head -c 300 gpurun_out/r2t_bench.json; echo
timeout 300 python -m pytest tests/test_gpu_accuracy.py tests/test_gpu_features.py tests/test_gpu_rshim.py tests/test_gpu_host_api.py -m gpu -x -q > gpurun_out/r2t_pytest.txt 2>&1; echo "pytest rc=$?"; tail -2 gpurun_out/r2t_pytest.txt
