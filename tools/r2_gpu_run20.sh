#!/bin/bash
# round-2 GPU session 20: the long parity file on the final tree (sessions 19 + 20 together = the whole GPU suite)
mkdir -p gpurun_out
timeout 125 python -m pytest tests/test_gpu_parity.py -m gpu -x -q > gpurun_out/r2v_pytest.txt 2>&1; echo "pytest rc=$?"; tail -2 gpurun_out/r2v_pytest.txt
