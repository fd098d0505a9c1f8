#!/bin/bash
# round-2 GPU session 16: the tree after the dead row-skipping plumbing was removed: smoke, accuracy / feature / shim tests, probe
mkdir -p gpurun_out
python __graft_entry__.py --smoke > gpurun_out/r2r_smoke.txt 2>&1; echo "smoke rc=$?"; tail -1 gpurun_out/r2r_smoke.txt
timeout 120 python tools/i8_probe.py 50000 1,0,0 > gpurun_out/r2r_probe.txt 2>&1; echo "probe rc=$?"; tail -1 gpurun_out/r2r_probe.txt
timeout 600 python -m pytest tests/test_gpu_accuracy.py tests/test_gpu_features.py tests/test_gpu_rshim.py tests/test_gpu_host_api.py -m gpu -x -q > gpurun_out/r2r_pytest.txt 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/r2r_pytest.txt
