#!/bin/bash
# round-2 GPU session 8: epilogue compaction (FP64 work only for the rows with a mask entry) against the every-row epilogue (dense7)
mkdir -p gpurun_out
L=genomicmateselectr_b200/lib
python __graft_entry__.py --smoke > gpurun_out/r2j_smoke.txt 2>&1; echo "smoke rc=$?"; tail -1 gpurun_out/r2j_smoke.txt
timeout 200 python tools/i8_probe.py 50000 1,0,0 0,0,0 > gpurun_out/r2j_probe_dig7.txt 2>&1; echo "probe7 rc=$?"; tail -2 gpurun_out/r2j_probe_dig7.txt
GMS_B200_LIB=$L/dense7/libgms_b200.so timeout 200 python tools/i8_probe.py 50000 1,0,0 > gpurun_out/r2j_probe_dense7.txt 2>&1; echo "dense7 rc=$?"; tail -1 gpurun_out/r2j_probe_dense7.txt
GMS_B200_LIB=$L/dig6/libgms_b200.so timeout 200 python tools/i8_probe.py 50000 1,0,0 > gpurun_out/r2j_probe_dig6.txt 2>&1; echo "probe6 rc=$?"; tail -1 gpurun_out/r2j_probe_dig6.txt
GMS_B200_LIB=$L/split7/libgms_b200.so timeout 90 python tools/i8_probe.py 50000 1,0,0 > gpurun_out/r2j_probe_split7.txt 2>&1; echo "split7 rc=$?"; tail -1 gpurun_out/r2j_probe_split7.txt
GMS_B200_LIB=$L/split6/libgms_b200.so timeout 90 python tools/i8_probe.py 50000 1,0,0 > gpurun_out/r2j_probe_split6.txt 2>&1; echo "split6 rc=$?"; tail -1 gpurun_out/r2j_probe_split6.txt
GMS_B200_LIB=$L/profile7/libgms_b200.so timeout 300 python tools/i8_probe.py 50000 1,0,0 1,0,1 1,0,32 1,0,64 1,0,2 > gpurun_out/r2j_probe_profile7.txt 2>&1; echo "profile7 rc=$?"; tail -5 gpurun_out/r2j_probe_profile7.txt
GMS_B200_LIB=$L/split7/libgms_b200.so timeout 300 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "synthetic_map_path or blocks_path or int8_tensor or cassava_scale_against" > gpurun_out/r2j_pytest_split7.txt 2>&1; echo "pytest split7 rc=$?"; tail -2 gpurun_out/r2j_pytest_split7.txt
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_accuracy.py tests/test_gpu_fuzz.py -m gpu -x -q > gpurun_out/r2j_pytest.txt 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/r2j_pytest.txt
