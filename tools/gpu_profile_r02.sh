#!/bin/bash
# Round-2 profiles for profiles/: launch lists (ncu, gpu__time_duration) of one iteration of the per-GPU shard of
# BASELINE configs 3, 4 and 5, and --set full captures of the row-wise kernels and the MSA column attention.
#   gpurun -- 'bash tools/gpu_profile_r02.sh r02'
TAG=${1:-r02}
OUT=gpurun_out
mkdir -p $OUT
for c in c3 c4 c5; do
  timeout 500 ncu --metrics gpu__time_duration.sum --clock-control none -c 2500 --csv \
      --log-file $OUT/${TAG}_launches_$c.csv python tools/config_bench.py $c > $OUT/${TAG}_ncu_$c.log 2>&1
  echo "ncu list $c exit $?"; wc -l $OUT/${TAG}_launches_$c.csv
done
# row-wise kernels at config 2 (bench timed region), MSA column / row attention at config 3
timeout 600 ncu --set full --clock-control none --import-source on --profile-from-start off \
    -k regex:'layernorm_kernel|embed_kernel|head_sample_kernel' -c 6 -f -o $OUT/${TAG}_rowwise \
    python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-other-configs > $OUT/${TAG}_ncu_rowwise.log 2>&1
echo "ncu rowwise exit $?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'msa_col_attention|msa_row_attention_tc' -s 4 -c 2 -f \
    -o $OUT/${TAG}_msa python tools/msa_profile.py > $OUT/${TAG}_ncu_msa.log 2>&1
echo "ncu msa exit $?"
timeout 600 ncu --set full --clock-control none --import-source on --profile-from-start off \
    -k regex:gemm_tcgen05 -s 8 -c 4 -f -o $OUT/${TAG}_gemm python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-other-configs > $OUT/${TAG}_ncu_gemm.log 2>&1
echo "ncu gemm exit $?"
ls -la $OUT | grep $TAG
