#!/bin/bash
# Same-box comparison of several builds of the library on the bench line (config 2 only, no CPU arm), two rounds:
#   gpurun -- 'bash tools/gpu_lib_abc.sh tag libA.so libB.so libC.so'
TAG=$1; shift
mkdir -p gpurun_out
for round in 1 2; do for lib in "$@"; do
  PGIBBS_LIB_PATH=$PWD/$lib timeout 400 python bench.py --steps 20 --warmup 3 --no-cpu-baseline --no-other-configs > gpurun_out/${TAG}_tmp.json 2> gpurun_out/${TAG}_tmp.err
  python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/${TAG}_tmp.json").read().strip().splitlines()[-1])
    print("$lib", "iters/s", round(d["value"],2), "ms", round(d["ms_per_step"],3), "clocks", d["clocks"]["sm_mhz"], {k:round(v["avg_launch_ms"],4) for k,v in d["roofline"]["per_kernel"].items()})
except Exception as e:
    print("$lib failed", e); print(open("gpurun_out/${TAG}_tmp.err").read()[-800:])
PY
done; done 2>&1 | tee gpurun_out/${TAG}_lib_abc.txt
