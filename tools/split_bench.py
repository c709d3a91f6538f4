"""Last-wave K-split of the residual GEMMs (GPU box): correctness (bit-identical repeat runs, value vs torch) and time
with PGIBBS_GEMM_SPLIT=1 vs 0 on the out-projection / FC2 shapes.
    python tools/split_bench.py > gpurun_out/split_bench.txt
"""
import os, subprocess, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

if len(sys.argv) > 1:
    import torch
    from protein_gibbs_sampler_b200.engine import op_gemm
    torch.manual_seed(0)
    for name, M, N, K in (("out", 16512, 1280, 1280), ("fc2", 16512, 1280, 5120), ("msa_out", 66048, 768, 768),
                          ("msa_fc2", 66048, 768, 3072), ("c4_fc2", 32896, 1280, 5120), ("small", 5000, 320, 1280)):
        A, B, bias = torch.randn(M, K) * 0.5, torch.randn(N, K) * 0.05, torch.randn(N)
        C0 = torch.randn(M, N)
        want = (A.half().cuda() @ B.half().cuda().t()).float().cpu() + bias + C0
        got, ms = op_gemm(A, B, bias, C=C0, epilogue=2, reps=30)
        got2 = op_gemm(A, B, bias, C=C0, epilogue=2)
        err = ((got - want).abs().max() / want.abs().max()).item()
        print("split=%s %-8s M%-6d N%-5d K%-5d %.4f ms %7.1f TFLOP/s rel err %.2e repeat-identical %s" % (
            os.environ.get("PGIBBS_GEMM_SPLIT", "1"), name, M, N, K, ms, 2.0 * M * N * K / (ms * 1e-3) / 1e12, err,
            bool((got == got2).all())), flush=True)
else:
    for split in ("1", "0"):
        env = dict(os.environ, PGIBBS_GEMM_SPLIT=split)
        subprocess.run([sys.executable, os.path.abspath(__file__), "run"], env=env, check=False)
