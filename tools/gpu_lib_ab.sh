#!/bin/bash
# Same-box A/B of two builds of the library on the bench line (config 2 only, no CPU arm):
#   gpurun -- 'bash tools/gpu_lib_ab.sh tag libA.so libB.so'   -> A B A B
TAG=$1; A=$2; B=$3
mkdir -p gpurun_out
for lib in $A $B $A $B; do
  PGIBBS_LIB_PATH=$PWD/$lib timeout 400 python bench.py --steps 20 --warmup 3 --no-cpu-baseline --no-other-configs > gpurun_out/${TAG}_tmp.json 2> gpurun_out/${TAG}_tmp.err
  python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/${TAG}_tmp.json").read().strip().splitlines()[-1])
    s=d["roofline"]["time_share_by_kernel"]
    print("$lib", "iters/s", round(d["value"],2), "ms", round(d["ms_per_step"],3), "e2e", round(d["e2e"]["value"],2), "attn share", s.get("attention"), "per-kernel ms", {k:round(v["avg_launch_ms"],4) for k,v in d["roofline"]["per_kernel"].items()})
except Exception as e:
    print("$lib failed", e); print(open("gpurun_out/${TAG}_tmp.err").read()[-800:])
PY
done 2>&1 | tee gpurun_out/${TAG}_lib_ab.txt
