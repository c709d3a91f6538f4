#!/bin/bash
# Final round-2 GPU pass: full GPU parity suite, smoke, the bench line, the ncu launch list of the bench's timed region
# (config 2) and the reference arm.   gpurun --timeout 2400 -- 'bash tools/gpu_final_r02.sh r02zz'
TAG=${1:-r02zz}
OUT=gpurun_out
mkdir -p $OUT
echo "== pytest -m gpu"
timeout 1200 python -m pytest tests -m gpu -x -q > $OUT/${TAG}_pytest_gpu.log 2>&1
echo "pytest exit $?"; tail -4 $OUT/${TAG}_pytest_gpu.log
echo "== smoke"
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -3
echo "== bench"
timeout 900 python bench.py --steps 20 --warmup 3 > $OUT/${TAG}_bench.json 2> $OUT/${TAG}_bench.err
echo "bench exit $?"; tail -c 400 $OUT/${TAG}_bench.json; echo
echo "== reference arm"
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > $OUT/${TAG}_ref.json 2> $OUT/${TAG}_ref.err
echo "ref exit $?"; cat $OUT/${TAG}_ref.json | cut -c1-600
echo "== ncu launch list (timed region of bench.py, 2 steps, config 2 only)"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off -c 600 --csv \
    --log-file $OUT/${TAG}_launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-other-configs > $OUT/${TAG}_ncu_bench.log 2>&1
echo "ncu list exit $?"; wc -l $OUT/${TAG}_launches.csv
if [ -n "$NCU_GEMM" ]; then
echo "== ncu --set full on the GEMMs of one layer"
timeout 600 ncu --set full --clock-control none --import-source on --profile-from-start off \
    -k regex:gemm_tcgen05 -s 8 -c 4 -o $OUT/${TAG}_gemm python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-other-configs \
    > $OUT/${TAG}_ncu_gemm.log 2>&1
echo "ncu full exit $?"; ls -la $OUT/${TAG}_gemm.ncu-rep
fi
