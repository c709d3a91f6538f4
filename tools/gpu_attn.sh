#!/bin/bash
# attention kernel A/B on one GPU box: fa (default), fa without the tail split, legacy; ncu of the fa kernel
OUT=gpurun_out
mkdir -p $OUT
for mode in ${MODES:-fa fa_notail}; do
  echo "== $mode"
  case $mode in
    fa) env="PGIBBS_ATTN=fa";;
    fa_notail) env="PGIBBS_ATTN=fa PGIBBS_ATTN_TAIL=0";;
    legacy) env="PGIBBS_ATTN=legacy";;
  esac
  env $env timeout 300 python tools/attn_bench.py > $OUT/attn_$mode.txt 2>&1
  echo "exit $?"; cat $OUT/attn_$mode.txt | tail -16
done
if [ -n "$NCU" ]; then
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:attention -c 6 -o $OUT/attn_fa \
      python tools/attn_one.py > $OUT/attn_ncu.log 2>&1
  echo "ncu exit $?"; tail -3 $OUT/attn_ncu.log
fi
