#!/bin/bash
# round-2 GPU session 12: seven digit planes as the product again, FP64 burst in the tensor-idle window (GMS_I8_LATE): full GPU
# suite, bench, ncu launch list of the bench command, one full ncu capture of the per-cross int8 kernel
mkdir -p gpurun_out
python __graft_entry__.py --smoke > gpurun_out/r2n_smoke.txt 2>&1; echo "smoke rc=$?"; tail -1 gpurun_out/r2n_smoke.txt
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r2n_pytest.txt 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/r2n_pytest.txt
timeout 600 python bench.py --steps 20 --warmup 5 > gpurun_out/r2n_bench.json 2> gpurun_out/r2n_bench.err; echo "bench rc=$?"; head -c 400 gpurun_out/r2n_bench.json; echo
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2n_launches.csv python bench.py --steps 2 --warmup 3 --modes auto --no-cpu-baseline --no-cached > gpurun_out/r2n_ncu_launches.log 2>&1; echo "ncu launches rc=$?"
timeout 400 ncu --set full --clock-control none --import-source on -k regex:i8_contract_kernel --launch-skip 5 --launch-count 1 -o gpurun_out/r2n_i8 -f python tools/i8_probe.py 50000 1,0,0 > gpurun_out/r2n_ncu.log 2>&1; echo "ncu full rc=$?"
