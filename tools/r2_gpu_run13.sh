#!/bin/bash
# round-2 GPU session 13: unit order chromosome group > trait > tile in the int8 kernel (operand re-reads from L2 instead of DRAM):
# probe timings, full GPU suite, bench, ncu launch list of the bench command, full ncu capture of the per-cross int8 kernel, large config
mkdir -p gpurun_out
python __graft_entry__.py --smoke > gpurun_out/r2o_smoke.txt 2>&1; echo "smoke rc=$?"; tail -1 gpurun_out/r2o_smoke.txt
timeout 120 python tools/i8_probe.py 50000 1,0,0 > gpurun_out/r2o_probe.txt 2>&1; echo "probe rc=$?"; tail -1 gpurun_out/r2o_probe.txt
PROBE_WORKLOAD=large timeout 300 python tools/i8_probe.py 20000 1,0,0 > gpurun_out/r2o_probe_large.txt 2>&1; echo "large rc=$?"; tail -1 gpurun_out/r2o_probe_large.txt
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r2o_pytest.txt 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/r2o_pytest.txt
timeout 600 python bench.py --steps 20 --warmup 5 > gpurun_out/r2o_bench.json 2> gpurun_out/r2o_bench.err; echo "bench rc=$?"; head -c 300 gpurun_out/r2o_bench.json; echo
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2o_launches.csv python bench.py --steps 2 --warmup 3 --modes auto --no-cpu-baseline --no-cached > gpurun_out/r2o_ncu_launches.log 2>&1; echo "ncu launches rc=$?"
timeout 400 ncu --set full --clock-control none --import-source on -k regex:i8_contract_kernel --launch-skip 5 --launch-count 1 -o gpurun_out/r2o_i8 -f python tools/i8_probe.py 50000 1,0,0 > gpurun_out/r2o_ncu.log 2>&1; echo "ncu full rc=$?"
timeout 400 python bench.py --workload large --crosses-per-gpu 20000 --steps 3 --warmup 3 --modes auto --no-cpu-baseline --no-cached > gpurun_out/r2o_bench_large.json 2> gpurun_out/r2o_bench_large.err; echo "bench large rc=$?"; head -c 300 gpurun_out/r2o_bench_large.json; echo
