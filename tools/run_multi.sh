#!/bin/bash
# usage: tools/run_multi.sh <ngpus> <tag> [bench args...]   -> gpurun_out/<tag>.json / .err
N=$1; TAG=$2; shift 2
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus $N "$@" > gpurun_out/$TAG.json 2> gpurun_out/$TAG.err
echo "$TAG rc=$?"; tail -2 gpurun_out/$TAG.err | cut -c1-300
python - <<PY
import json
try:
    d=json.load(open("gpurun_out/$TAG.json"))
    print("$TAG", "value", round(d["value"]), "e2e", round(d["e2e"]["value"]), "ms/step", round(d["ms_per_step"],2), d["stages_ms"], d["config"]["grid"], d["scaling"])
except Exception as e:
    print("$TAG: no json", e)
PY
