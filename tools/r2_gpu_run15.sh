#!/bin/bash
# round-2 GPU session 15: verify accumulate with trait-contiguous staging on top of the group-outer kernel: full GPU suite, bench, launch list, ncu capture, large config
mkdir -p gpurun_out
python __graft_entry__.py --smoke > gpurun_out/r2q_smoke.txt 2>&1; echo "smoke rc=$?"; tail -1 gpurun_out/r2q_smoke.txt
timeout 120 python tools/i8_probe.py 50000 1,0,0 > gpurun_out/r2q_probe.txt 2>&1; echo "probe rc=$?"; tail -1 gpurun_out/r2q_probe.txt
PROBE_WORKLOAD=large timeout 300 python tools/i8_probe.py 20000 1,0,0 > gpurun_out/r2q_probe_large.txt 2>&1; echo "large rc=$?"; tail -1 gpurun_out/r2q_probe_large.txt
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r2q_pytest.txt 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/r2q_pytest.txt
timeout 600 python bench.py --steps 20 --warmup 5 > gpurun_out/r2q_bench.json 2> gpurun_out/r2q_bench.err; echo "bench rc=$?"; head -c 300 gpurun_out/r2q_bench.json; echo
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2q_launches.csv python bench.py --steps 2 --warmup 3 --modes auto --no-cpu-baseline --no-cached > gpurun_out/r2q_ncu_launches.log 2>&1; echo "ncu launches rc=$?"
timeout 400 ncu --set full --clock-control none --import-source on -k regex:i8_contract_kernel --launch-skip 5 --launch-count 1 -o gpurun_out/r2q_i8 -f python tools/i8_probe.py 50000 1,0,0 > gpurun_out/r2q_ncu.log 2>&1; echo "ncu full rc=$?"
timeout 400 python bench.py --workload large --crosses-per-gpu 20000 --steps 3 --warmup 3 --modes auto --no-cpu-baseline --no-cached > gpurun_out/r2q_bench_large.json 2> gpurun_out/r2q_bench_large.err; echo "bench large rc=$?"; head -c 300 gpurun_out/r2q_bench_large.json; echo
