"""Epilogue cost on config 3's out-projection shape (MSA-1b: M = 16 x 32 x 129 rows, N = K = 768) and FC2 (K = 3072): the
fp32 reduce-add epilogue (2), a plain fp32 store (5) and an fp16 store (0).   python tools/epi_bench_c3.py"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from protein_gibbs_sampler_b200.engine import op_gemm

torch.manual_seed(0)
M = 16 * 32 * 129
for name, N, K in (("out", 768, 768), ("fc2", 768, 3072)):
    A, B, bias = torch.randn(M, K) * 0.5, torch.randn(N, K) * 0.05, torch.randn(N)
    C0 = torch.randn(M, N)
    for epi in (2, 5, 0):
        for rep in range(2):
            got, ms = op_gemm(A, B, bias, C=C0 if epi == 2 else None, epilogue=epi, block_n=256, cta_group=2, reps=30)
        print("%-4s M%d N%-5d K%-5d epi%d  %.4f ms  %7.1f TFLOP/s" % (name, M, N, K, epi, ms, 2.0 * M * N * K / (ms * 1e-3) / 1e12), flush=True)
