#!/bin/bash
mkdir -p gpurun_out
timeout 300 python tools/dbg_blocks.py > gpurun_out/r2i_dbg.txt 2>&1; echo "rc=$?"; tail -40 gpurun_out/r2i_dbg.txt
