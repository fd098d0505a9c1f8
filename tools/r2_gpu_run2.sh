#!/bin/bash
# round-2 GPU session 2: drain-then-release epilogue; 7 vs 6 digit planes; where the time goes (profiling build); tests; bench; ncu
mkdir -p gpurun_out
L=genomicmateselectr_b200/lib
python __graft_entry__.py --smoke > gpurun_out/r2d_smoke.txt 2>&1; echo "smoke rc=$?"; tail -1 gpurun_out/r2d_smoke.txt
timeout 300 python tools/i8_probe.py 50000 1,1,0 1,0,0 0,1,0 > gpurun_out/r2d_probe_dig7.txt 2>&1; echo "probe7 rc=$?"; cat gpurun_out/r2d_probe_dig7.txt | tail -4
GMS_B200_LIB=$L/dig6/libgms_b200.so timeout 300 python tools/i8_probe.py 50000 1,1,0 1,0,0 > gpurun_out/r2d_probe_dig6.txt 2>&1; echo "probe6 rc=$?"; tail -3 gpurun_out/r2d_probe_dig6.txt
GMS_B200_LIB=$L/profile7/libgms_b200.so timeout 400 python tools/i8_probe.py 50000 1,1,0 1,1,1 1,1,32 1,1,2 1,1,4 1,1,8 1,1,16 1,1,24 1,1,3 > gpurun_out/r2d_probe_profile7.txt 2>&1; echo "profile7 rc=$?"; tail -10 gpurun_out/r2d_probe_profile7.txt
GMS_B200_LIB=$L/profile6/libgms_b200.so timeout 300 python tools/i8_probe.py 50000 1,1,0 1,1,1 1,1,2 > gpurun_out/r2d_probe_profile6.txt 2>&1; echo "profile6 rc=$?"; tail -4 gpurun_out/r2d_probe_profile6.txt
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/r2d_pytest.txt 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/r2d_pytest.txt
timeout 600 python bench.py --steps 5 --warmup 3 --modes auto,fp64,map_scan > gpurun_out/r2d_bench.json 2> gpurun_out/r2d_bench.err; echo "bench rc=$?"
python - <<'PY'
import json
try:
    d = json.load(open("gpurun_out/r2d_bench.json"))
    print("value", d["value"], "e2e", d["e2e"]["value"], "kernel_ms", d["roofline"]["kernel_ms"], "frac", d["roofline"]["frac"], "fallback", d["fallback"], "stages", d["stages_ms"])
    print({k: (v["max_rel_err_variances_strict"], v["max_rel_err_all_guarded"]) for k, v in d["parity"].items()})
except Exception as e:
    print("no bench json", e)
PY
timeout 600 ncu --set full --clock-control none --import-source on -k regex:i8_contract_kernel --launch-skip 5 --launch-count 1 -o gpurun_out/r2d_i8 -f python tools/i8_probe.py 50000 1,1,0 > gpurun_out/r2d_ncu.log 2>&1; echo "ncu rc=$?"
