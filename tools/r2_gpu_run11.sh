#!/bin/bash
# round-2 GPU session 11: six digit planes by default, late FP64 burst for T <= 10: full GPU suite, bench, large-config probe
mkdir -p gpurun_out
L=genomicmateselectr_b200/lib
python __graft_entry__.py --smoke > gpurun_out/r2m_smoke.txt 2>&1; echo "smoke rc=$?"; tail -1 gpurun_out/r2m_smoke.txt
timeout 120 python tools/i8_probe.py 50000 1,0,0 > gpurun_out/r2m_probe_dig6.txt 2>&1; echo "probe6 rc=$?"; tail -1 gpurun_out/r2m_probe_dig6.txt
GMS_B200_LIB=$L/dig7/libgms_b200.so timeout 120 python tools/i8_probe.py 50000 1,0,0 > gpurun_out/r2m_probe_dig7.txt 2>&1; echo "probe7 rc=$?"; tail -1 gpurun_out/r2m_probe_dig7.txt
PROBE_WORKLOAD=large timeout 300 python tools/i8_probe.py 20000 1,0,0 > gpurun_out/r2m_probe_large.txt 2>&1; echo "large rc=$?"; tail -1 gpurun_out/r2m_probe_large.txt
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r2m_pytest.txt 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/r2m_pytest.txt
timeout 600 python bench.py --steps 20 --warmup 5 > gpurun_out/r2m_bench.json 2> gpurun_out/r2m_bench.err; echo "bench rc=$?"; tail -c 600 gpurun_out/r2m_bench.json
