#!/bin/bash
# A/B of one environment switch on the bench line:  gpurun -- 'bash tools/gpu_ab.sh PGIBBS_ZIGZAG tag [steps]'
VAR=$1; TAG=${2:-ab}; STEPS=${3:-10}
mkdir -p gpurun_out
for v in 0 1 0 1; do
env $VAR=$v timeout 600 python bench.py --steps $STEPS --warmup 3 --no-cpu-baseline > gpurun_out/${TAG}_$v.json 2> gpurun_out/${TAG}_$v.err
python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/${TAG}_$v.json").read().strip().splitlines()[-1])
    print("$VAR=$v", round(d["value"],2), round(d["ms_per_step"],3), round(d["e2e"]["value"],2), d["roofline"]["time_share_by_kernel"])
except Exception as e:
    print("$VAR=$v failed", e); print(open("gpurun_out/${TAG}_$v.err").read()[-1500:])
PY
done
