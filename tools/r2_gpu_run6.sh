#!/bin/bash
# round-2 GPU session 6: 64-blocks dealt evenly to the row tiles (common.cuh) against the lower block triangle (tri), 6 / 7 digit planes
mkdir -p gpurun_out
L=genomicmateselectr_b200/lib
python __graft_entry__.py --smoke > gpurun_out/r2h_smoke.txt 2>&1; echo "smoke rc=$?"; tail -1 gpurun_out/r2h_smoke.txt
timeout 200 python tools/i8_probe.py 50000 1,0,0 > gpurun_out/r2h_probe_dig7.txt 2>&1; echo "probe7 rc=$?"; tail -1 gpurun_out/r2h_probe_dig7.txt
GMS_B200_LIB=$L/tri/libgms_b200.so timeout 200 python tools/i8_probe.py 50000 1,0,0 > gpurun_out/r2h_probe_tri.txt 2>&1; echo "tri rc=$?"; tail -1 gpurun_out/r2h_probe_tri.txt
GMS_B200_LIB=$L/dig6/libgms_b200.so timeout 200 python tools/i8_probe.py 50000 1,0,0 > gpurun_out/r2h_probe_dig6.txt 2>&1; echo "probe6 rc=$?"; tail -1 gpurun_out/r2h_probe_dig6.txt
GMS_B200_LIB=$L/dig6np3/libgms_b200.so timeout 200 python tools/i8_probe.py 50000 1,0,0 > gpurun_out/r2h_probe_dig6np3.txt 2>&1; echo "probe6np3 rc=$?"; tail -1 gpurun_out/r2h_probe_dig6np3.txt
GMS_B200_LIB=$L/profile7/libgms_b200.so timeout 300 python tools/i8_probe.py 50000 1,0,0 1,0,1 1,0,32 1,0,64 1,0,2 > gpurun_out/r2h_probe_profile7.txt 2>&1; echo "profile7 rc=$?"; tail -5 gpurun_out/r2h_probe_profile7.txt
timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_gpu_accuracy.py -m gpu -x -q > gpurun_out/r2h_pytest.txt 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/r2h_pytest.txt
