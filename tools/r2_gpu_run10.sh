#!/bin/bash
# round-2 GPU session 10: FP64 burst of the previous unit interleaved with the drain (default; fnp2: two partial sums per trait)
# against the burst before the drain (late7)
mkdir -p gpurun_out
L=genomicmateselectr_b200/lib
python __graft_entry__.py --smoke > gpurun_out/r2l_smoke.txt 2>&1; echo "smoke rc=$?"; tail -1 gpurun_out/r2l_smoke.txt
timeout 120 python tools/i8_probe.py 50000 1,0,0 > gpurun_out/r2l_probe_dig7.txt 2>&1; echo "probe7 rc=$?"; tail -1 gpurun_out/r2l_probe_dig7.txt
GMS_B200_LIB=$L/fnp2/libgms_b200.so timeout 120 python tools/i8_probe.py 50000 1,0,0 > gpurun_out/r2l_probe_fnp2.txt 2>&1; echo "fnp2 rc=$?"; tail -1 gpurun_out/r2l_probe_fnp2.txt
GMS_B200_LIB=$L/late7/libgms_b200.so timeout 120 python tools/i8_probe.py 50000 1,0,0 > gpurun_out/r2l_probe_late7.txt 2>&1; echo "late7 rc=$?"; tail -1 gpurun_out/r2l_probe_late7.txt
GMS_B200_LIB=$L/dig6/libgms_b200.so timeout 120 python tools/i8_probe.py 50000 1,0,0 > gpurun_out/r2l_probe_dig6.txt 2>&1; echo "probe6 rc=$?"; tail -1 gpurun_out/r2l_probe_dig6.txt
GMS_B200_LIB=$L/profile7/libgms_b200.so timeout 200 python tools/i8_probe.py 50000 1,0,0 1,0,1 1,0,32 1,0,64 > gpurun_out/r2l_probe_profile7.txt 2>&1; echo "profile7 rc=$?"; tail -4 gpurun_out/r2l_probe_profile7.txt
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q > gpurun_out/r2l_pytest.txt 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/r2l_pytest.txt
