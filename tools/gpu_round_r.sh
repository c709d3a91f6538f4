#!/bin/bash
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -x -q -m gpu > gpurun_out/r01r_tests.txt 2>&1
tail -3 gpurun_out/r01r_tests.txt
for v in 0 1; do
echo "PGIBBS_PDL=$v"
PGIBBS_PDL=$v timeout 600 python tools/config_bench.py c1 2>&1 | grep -v warning
done > gpurun_out/r01r_latency_pdl.txt
cat gpurun_out/r01r_latency_pdl.txt
timeout 600 python bench.py --steps 10 --warmup 3 > gpurun_out/r01r_bench.json 2> gpurun_out/r01r_bench.err
python -c "
import json
d=json.loads(open('gpurun_out/r01r_bench.json').read().strip().splitlines()[-1])
print(round(d['value'],2), round(d['e2e']['value'],2), d['gpu_launches'])"
