"""GEMM tiling sweep on the BASELINE config-2 shapes (GPU box): correctness vs torch + time per launch.
    python tools/gemm_bench.py > gpurun_out/gemm_bench.txt
"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from protein_gibbs_sampler_b200.engine import op_gemm

torch.manual_seed(0)
M = int(os.environ.get("GEMM_M", 16512))
shapes = [("qkv", 3840, 1280, 0), ("out", 1280, 1280, 2), ("fc1", 5120, 1280, 1), ("fc2", 1280, 5120, 2)]
small = [(300, 320, 320), (129, 336, 128), (257, 512, 64), (1000, 1280, 1280)]
print("== correctness, small/ragged shapes, CTA pairs")
for (m, n, k) in small:
    A, B, bias = torch.randn(m, k) * 0.5, torch.randn(n, k) * 0.5, torch.randn(n)
    want = A.half().float() @ B.half().float().t() + bias
    for bn in (128, 192, 256):
        got = op_gemm(A, B, bias, epilogue=5, block_n=bn, cta_group=2)
        err = ((got - want).abs().max() / want.abs().max()).item()
        print("M%d N%d K%d bn%d cg2 rel err %.2e %s" % (m, n, k, bn, err, "ok" if err < 2e-5 else "FAIL"))
print("== sweep M=%d" % M)
for name, N, K, epi in shapes:
    A, B, bias = torch.randn(M, K) * 0.5, torch.randn(N, K) * 0.05, torch.randn(N)
    want = (A.half().cuda() @ B.half().cuda().t()).float().cpu() + bias
    if epi == 1:
        want = torch.nn.functional.gelu(want)
    C0 = torch.randn(M, N) if epi == 2 else None
    if epi == 2:
        want = want + C0
    # cuBLAS on the same shape, same process / thermal state: the number our kernel is measured against
    a16, b16 = A.half().cuda(), B.half().cuda()
    for _ in range(3):
        a16 @ b16.t()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(20):
        a16 @ b16.t()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 20
    print("%-4s N%-5d K%-5d cuBLAS fp16 (no epilogue)  %.3f ms  %7.1f TFLOP/s" % (name, N, K, ms, 2.0 * M * N * K / (ms * 1e-3) / 1e12))
    del a16, b16
    for cg in (1, 2):
        for bn in (128, 192, 256):
            got, ms = op_gemm(A, B, bias, C=C0, epilogue=epi, block_n=bn, cta_group=cg, reps=20)
            err = ((got - want).abs().max() / want.abs().max()).item()
            tf = 2.0 * M * N * K / (ms * 1e-3) / 1e12
            print("%-4s N%-5d K%-5d cg%d bn%-3d  %.3f ms  %7.1f TFLOP/s  rel err %.2e" % (name, N, K, cg, bn, ms, tf, err),
                  flush=True)
