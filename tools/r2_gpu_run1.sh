#!/bin/bash
# round-2 GPU session: smoke, int8 kernel variants (sire order / row skipping), GPU tests, bench
mkdir -p gpurun_out
python __graft_entry__.py --smoke > gpurun_out/r2c_smoke.txt 2>&1; echo "smoke rc=$?"
tail -2 gpurun_out/r2c_smoke.txt
for v in "1 1" "0 0" "1 0" "0 1"; do
  set -- $v
  GMS_I8_SORT=$1 GMS_I8_SKIP=$2 timeout 300 python tools/profile_modes.py 50000 int8_tensor > gpurun_out/r2c_modes_s$1_k$2.txt 2>&1
  echo "sort=$1 skip=$2 rc=$?"; grep int8_tensor gpurun_out/r2c_modes_s$1_k$2.txt | cut -c1-200
done
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/r2c_pytest.txt 2>&1; echo "pytest rc=$?"
tail -5 gpurun_out/r2c_pytest.txt
timeout 600 python bench.py --steps 5 --warmup 3 --modes auto,fp64,map_scan > gpurun_out/r2c_bench.json 2> gpurun_out/r2c_bench.err; echo "bench rc=$?"
python - <<'PY'
import json
try:
    d = json.load(open("gpurun_out/r2c_bench.json"))
    print("value", d["value"], "e2e", d["e2e"]["value"], "kernel_ms", d["roofline"]["kernel_ms"], "frac", d["roofline"]["frac"], "parity", d["parity"], "fallback", d["fallback"], "stages", d["stages_ms"])
except Exception as e:
    print("no bench json", e)
PY
