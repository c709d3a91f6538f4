#!/bin/bash
# new scoring / CLI tests + epilogue bench
mkdir -p gpurun_out
timeout 900 python -m pytest tests -x -q -m gpu -k "scoring or likelihood or from_fasta or loglik or log_lik" > gpurun_out/r01l_tests.txt 2>&1
tail -5 gpurun_out/r01l_tests.txt
timeout 300 python tools/epi_bench.py > gpurun_out/r01l_epi_bench.txt 2>&1
cat gpurun_out/r01l_epi_bench.txt
