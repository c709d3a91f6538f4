#!/bin/bash
# same-box A/B of two library builds on config 4 (ESM-2 650M, 64 x L512):  bash tools/gpu_c4_ab.sh libA.so libB.so
for lib in $1 $2 $1 $2; do
  PGIBBS_LIB_PATH=$PWD/$lib timeout 400 python tools/config_bench.py c4 2>&1 | grep -v warning | tail -3 | sed "s|^|$lib  |"
done
