#!/bin/bash
mkdir -p gpurun_out
timeout 600 python tools/split_bench.py > gpurun_out/r01m_split_bench.txt 2>&1
cat gpurun_out/r01m_split_bench.txt
timeout 900 python -m pytest tests/test_gpu.py -x -q -k "gemm_operator or forward_logits or full_size or golden" > gpurun_out/r01m_tests.txt 2>&1
tail -5 gpurun_out/r01m_tests.txt
PGIBBS_GEMM_SPLIT=0 timeout 600 python bench.py --steps 10 --warmup 3 > gpurun_out/r01m_bench_nosplit.json 2> gpurun_out/r01m_bench_nosplit.err
timeout 600 python bench.py --steps 10 --warmup 3 > gpurun_out/r01m_bench_split.json 2> gpurun_out/r01m_bench_split.err
PGIBBS_GEMM_SPLIT=0 timeout 600 python bench.py --steps 10 --warmup 3 > gpurun_out/r01m_bench_nosplit2.json 2> gpurun_out/r01m_bench_nosplit2.err
timeout 600 python bench.py --steps 10 --warmup 3 > gpurun_out/r01m_bench_split2.json 2> gpurun_out/r01m_bench_split2.err
python - <<'PY'
import json
for n in ("nosplit","split","nosplit2","split2"):
    try:
        d=json.loads(open("gpurun_out/r01m_bench_%s.json"%n).read().strip().splitlines()[-1])
        print(n, round(d["value"],2), round(d["ms_per_step"],3), d["roofline"]["time_share_by_kernel"])
    except Exception as e: print(n, "failed", e)
PY
