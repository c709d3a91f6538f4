#!/bin/bash
mkdir -p gpurun_out
for v in 0 1; do
echo "PGIBBS_PDL=$v"
PGIBBS_PDL=$v timeout 600 python tools/config_bench.py c1 2>&1 | grep -v warning
done > gpurun_out/r01q_latency_pdl.txt
cat gpurun_out/r01q_latency_pdl.txt
