#!/bin/bash
# round-2 GPU session 4: one dense FP64 burst per unit
mkdir -p gpurun_out
L=genomicmateselectr_b200/lib
python __graft_entry__.py --smoke > gpurun_out/r2f_smoke.txt 2>&1; echo "smoke rc=$?"; tail -1 gpurun_out/r2f_smoke.txt
timeout 300 python tools/i8_probe.py 50000 0,0,0 1,0,0 > gpurun_out/r2f_probe_dig7.txt 2>&1; echo "probe7 rc=$?"; tail -2 gpurun_out/r2f_probe_dig7.txt
GMS_B200_LIB=$L/dig6/libgms_b200.so timeout 300 python tools/i8_probe.py 50000 0,0,0 > gpurun_out/r2f_probe_dig6.txt 2>&1; echo "probe6 rc=$?"; tail -1 gpurun_out/r2f_probe_dig6.txt
GMS_B200_LIB=$L/profile7/libgms_b200.so timeout 400 python tools/i8_probe.py 50000 0,0,0 0,0,1 0,0,32 0,0,64 0,0,2 > gpurun_out/r2f_probe_profile7.txt 2>&1; echo "profile7 rc=$?"; tail -5 gpurun_out/r2f_probe_profile7.txt
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_accuracy.py -m gpu -x -q > gpurun_out/r2f_pytest.txt 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/r2f_pytest.txt
