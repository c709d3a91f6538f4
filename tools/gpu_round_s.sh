#!/bin/bash
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -x -q -m gpu > gpurun_out/r01s_tests.txt 2>&1
tail -8 gpurun_out/r01s_tests.txt
for v in 0 1; do
echo "PGIBBS_GRAPH=$v"
PGIBBS_GRAPH=$v timeout 600 python tools/config_bench.py c1 2>&1 | grep -v warning
done > gpurun_out/r01s_latency_graph.txt
cat gpurun_out/r01s_latency_graph.txt
for v in 0 1; do
PGIBBS_GRAPH=$v timeout 600 python bench.py --steps 10 --warmup 3 > gpurun_out/r01s_bench_graph$v.json 2> gpurun_out/r01s_bench_graph$v.err
python -c "
import json
d=json.loads(open('gpurun_out/r01s_bench_graph$v.json').read().strip().splitlines()[-1])
print('graph=$v', round(d['value'],2), round(d['e2e']['value'],2), d['gpu_launches'])" || tail -5 gpurun_out/r01s_bench_graph$v.err
done
