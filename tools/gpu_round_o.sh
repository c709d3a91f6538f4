#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu.py -x -q -k "forward_logits or golden" > gpurun_out/r01o_tests.txt 2>&1
tail -3 gpurun_out/r01o_tests.txt
PGIBBS_LN_FUSE=0 timeout 600 python bench.py --steps 10 --warmup 3 > gpurun_out/r01o_bench_nofuse.json 2> gpurun_out/r01o_bench_nofuse.err
timeout 600 python bench.py --steps 10 --warmup 3 > gpurun_out/r01o_bench_fuse.json 2> gpurun_out/r01o_bench_fuse.err
python - <<'PY'
import json
for n in ("nofuse","fuse"):
    try:
        d=json.loads(open("gpurun_out/r01o_bench_%s.json"%n).read().strip().splitlines()[-1])
        print(n, round(d["value"],2), round(d["ms_per_step"],3), d["roofline"]["time_share_by_kernel"])
    except Exception as e:
        print(n, "failed", e); print(open("gpurun_out/r01o_bench_%s.err"%n).read()[-1500:])
PY
