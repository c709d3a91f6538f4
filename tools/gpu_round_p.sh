#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu.py -x -q -k "forward_logits or golden or full_size or invariants or attention_operator or gemm_operator" > gpurun_out/r01p_tests.txt 2>&1
tail -3 gpurun_out/r01p_tests.txt
for v in 0 1 0 1; do
PGIBBS_PDL=$v timeout 600 python bench.py --steps 10 --warmup 3 > gpurun_out/r01p_bench_pdl$v.json 2> gpurun_out/r01p_bench_pdl$v.err
python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/r01p_bench_pdl$v.json").read().strip().splitlines()[-1])
    print("pdl=$v", round(d["value"],2), round(d["ms_per_step"],3), round(d["e2e"]["value"],2), d["roofline"]["time_share_by_kernel"])
except Exception as e:
    print("pdl=$v failed", e); print(open("gpurun_out/r01p_bench_pdl$v.err").read()[-1500:])
PY
done
