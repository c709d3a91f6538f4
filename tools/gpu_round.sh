#!/bin/bash
# One gpurun call: GPU parity tests, the bench line, the ncu launch list of the bench's timed region and one
# `--set full` capture of the dominant GEMM.  Everything lands in gpurun_out/.
#   gpurun --timeout 1500 -- 'bash tools/gpu_round.sh [tag]'
TAG=${1:-r01}
OUT=gpurun_out
mkdir -p $OUT
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw,memory.total --format=csv > $OUT/${TAG}_gpu.txt 2>&1

echo "== pytest -m gpu"
timeout 900 python -m pytest tests -m gpu -x -q > $OUT/${TAG}_pytest_gpu.log 2>&1
echo "pytest exit $?"; tail -5 $OUT/${TAG}_pytest_gpu.log

echo "== bench"
timeout 600 python bench.py --steps 20 --warmup 3 > $OUT/${TAG}_bench.json 2> $OUT/${TAG}_bench.err
echo "bench exit $?"; cat $OUT/${TAG}_bench.json; tail -3 $OUT/${TAG}_bench.err

echo "== ncu launch list (timed region of bench.py, 2 steps)"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off -c 600 --csv \
    --log-file $OUT/${TAG}_launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > $OUT/${TAG}_ncu_bench.log 2>&1
echo "ncu list exit $?"; wc -l $OUT/${TAG}_launches.csv

echo "== ncu --set full on the GEMMs of one layer"
timeout 600 ncu --set full --clock-control none --import-source on --profile-from-start off \
    -k regex:gemm_tcgen05 -s 8 -c 4 -o $OUT/${TAG}_gemm python bench.py --steps 1 --warmup 3 --no-cpu-baseline \
    > $OUT/${TAG}_ncu_gemm.log 2>&1
echo "ncu full exit $?"; ls -la $OUT

echo "== per-config table"
timeout 900 python tools/config_bench.py c1 c2 c3 c4 c5 2>&1 | grep -v warning > $OUT/${TAG}_config_bench.txt
cat $OUT/${TAG}_config_bench.txt
