#!/bin/bash
# round-2 GPU session 14: A/B on one box - unit order chromosome > trait > tile (default, old slab values preloaded at T <= 5)
# against trait > tile (lib/traitouter), cassava and large geometry; parity file
mkdir -p gpurun_out
L=genomicmateselectr_b200/lib
python __graft_entry__.py --smoke > gpurun_out/r2p_smoke.txt 2>&1; echo "smoke rc=$?"; tail -1 gpurun_out/r2p_smoke.txt
for rep in 1 2; do
timeout 120 python tools/i8_probe.py 50000 1,0,0 > gpurun_out/r2p_probe_group_$rep.txt 2>&1; echo "group rc=$?"; tail -1 gpurun_out/r2p_probe_group_$rep.txt
GMS_B200_LIB=$L/traitouter/libgms_b200.so timeout 120 python tools/i8_probe.py 50000 1,0,0 > gpurun_out/r2p_probe_trait_$rep.txt 2>&1; echo "trait rc=$?"; tail -1 gpurun_out/r2p_probe_trait_$rep.txt
done
PROBE_WORKLOAD=large timeout 300 python tools/i8_probe.py 20000 1,0,0 > gpurun_out/r2p_probe_large_group.txt 2>&1; echo "large group rc=$?"; tail -1 gpurun_out/r2p_probe_large_group.txt
GMS_B200_LIB=$L/traitouter/libgms_b200.so PROBE_WORKLOAD=large timeout 300 python tools/i8_probe.py 20000 1,0,0 > gpurun_out/r2p_probe_large_trait.txt 2>&1; echo "large trait rc=$?"; tail -1 gpurun_out/r2p_probe_large_trait.txt
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q > gpurun_out/r2p_pytest.txt 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/r2p_pytest.txt
