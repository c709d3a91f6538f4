#!/bin/bash
# attention kernel A/B builds on one GPU box: gpurun -- 'bash tools/gpu_attn_ab.sh tag lib1.so lib2.so ...'
# each library: operator parity test (tests/test_gpu.py -k attention_operator) and tools/attn_time.py twice
TAG=$1; shift
mkdir -p gpurun_out
for lib in "$@"; do
  echo "== $lib"
  PGIBBS_LIB_PATH=$PWD/$lib timeout 300 python -m pytest tests/test_gpu.py -q -x -m gpu -k "attention_operator or attention_operator_growing" 2>&1 | tail -3
  for r in 1 2; do PGIBBS_LIB_PATH=$PWD/$lib timeout 120 python tools/attn_time.py; done
done 2>&1 | tee gpurun_out/${TAG}_attn_ab.txt
