#!/bin/bash
# Same-box comparison of one build under several values of an environment switch, on the bench line (config 2 only, no
# CPU arm), two rounds:   gpurun -- 'bash tools/gpu_env_ab.sh tag PGIBBS_EPI_DIRECT 0 1 3 7'
TAG=$1; VAR=$2; shift 2
mkdir -p gpurun_out
for round in 1 2; do for val in "$@"; do
  env $VAR=$val timeout 400 python bench.py --steps 20 --warmup 3 --no-cpu-baseline --no-other-configs > gpurun_out/${TAG}_tmp.json 2> gpurun_out/${TAG}_tmp.err
  python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/${TAG}_tmp.json").read().strip().splitlines()[-1])
    print("$VAR=$val", "iters/s", round(d["value"],2), "ms", round(d["ms_per_step"],3), "e2e", round(d["e2e"]["value"],2), {k:round(v["avg_launch_ms"],4) for k,v in d["roofline"]["per_kernel"].items()})
except Exception as e:
    print("$VAR=$val failed", e); print(open("gpurun_out/${TAG}_tmp.err").read()[-800:])
PY
done; done 2>&1 | tee gpurun_out/${TAG}_env_ab.txt
