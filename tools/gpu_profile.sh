#!/bin/bash
# Profiles of the bench's timed region for profiles/: ncu launch list (2 steps), --set full of one layer's GEMMs,
# of the attention kernel and (config 3) of the MSA row-attention kernel.   gpurun -- 'bash tools/gpu_profile.sh tag'
TAG=${1:-r01}
OUT=gpurun_out
mkdir -p $OUT
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off -c 700 --csv \
    --log-file $OUT/${TAG}_launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > $OUT/${TAG}_ncu_bench.log 2>&1
echo "ncu list exit $?"; wc -l $OUT/${TAG}_launches.csv
timeout 600 ncu --set full --clock-control none --import-source on --profile-from-start off \
    -k regex:gemm_tcgen05 -s 8 -c 4 -o $OUT/${TAG}_gemm python bench.py --steps 1 --warmup 3 --no-cpu-baseline > $OUT/${TAG}_ncu_gemm.log 2>&1
echo "ncu gemm exit $?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:attention_fa -c 2 -o $OUT/${TAG}_attn \
    python tools/attn_one.py > $OUT/${TAG}_ncu_attn.log 2>&1
echo "ncu attn exit $?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:msa_row_attention_tc -s 2 -c 1 -o $OUT/${TAG}_msa_row \
    python tools/msa_profile.py > $OUT/${TAG}_ncu_msa.log 2>&1
echo "ncu msa exit $?"; ls -la $OUT | grep $TAG
