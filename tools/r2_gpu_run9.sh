#!/bin/bash
# round-2 GPU session 9: FP64 burst of unit u at the top of iteration u + 1 (tensor pipe idle) against the immediate burst (early7)
mkdir -p gpurun_out
L=genomicmateselectr_b200/lib
python __graft_entry__.py --smoke > gpurun_out/r2k_smoke.txt 2>&1; echo "smoke rc=$?"; tail -1 gpurun_out/r2k_smoke.txt
timeout 120 python tools/i8_probe.py 50000 1,0,0 > gpurun_out/r2k_probe_dig7.txt 2>&1; echo "probe7 rc=$?"; tail -1 gpurun_out/r2k_probe_dig7.txt
GMS_B200_LIB=$L/early7/libgms_b200.so timeout 120 python tools/i8_probe.py 50000 1,0,0 > gpurun_out/r2k_probe_early7.txt 2>&1; echo "early7 rc=$?"; tail -1 gpurun_out/r2k_probe_early7.txt
GMS_B200_LIB=$L/dig6/libgms_b200.so timeout 120 python tools/i8_probe.py 50000 1,0,0 > gpurun_out/r2k_probe_dig6.txt 2>&1; echo "probe6 rc=$?"; tail -1 gpurun_out/r2k_probe_dig6.txt
GMS_B200_LIB=$L/profile7/libgms_b200.so timeout 200 python tools/i8_probe.py 50000 1,0,0 1,0,1 1,0,32 1,0,64 > gpurun_out/r2k_probe_profile7.txt 2>&1; echo "profile7 rc=$?"; tail -4 gpurun_out/r2k_probe_profile7.txt
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q > gpurun_out/r2k_pytest.txt 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/r2k_pytest.txt
