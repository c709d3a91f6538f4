"""Epilogue cost on the out-projection / FC2 shapes (GPU box): same GEMM with the fp32 reduce-add epilogue (2), a plain
fp32 store (5) and an fp16 store (0).
    python tools/epi_bench.py > gpurun_out/epi_bench.txt
"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from protein_gibbs_sampler_b200.engine import op_gemm

torch.manual_seed(0)
M = int(os.environ.get("GEMM_M", 16512))
for name, N, K in (("out", 1280, 1280), ("fc2", 1280, 5120), ("qkv", 3840, 1280)):
    A, B, bias = torch.randn(M, K) * 0.5, torch.randn(N, K) * 0.05, torch.randn(N)
    C0 = torch.randn(M, N)
    for epi in (2, 5, 0):
        for rep in range(2):
            got, ms = op_gemm(A, B, bias, C=C0 if epi == 2 else None, epilogue=epi, block_n=256, cta_group=2, reps=30)
        print("%-4s N%-5d K%-5d epi%d  %.4f ms  %7.1f TFLOP/s" % (name, N, K, epi, ms, 2.0 * M * N * K / (ms * 1e-3) / 1e12),
              flush=True)
