"""Attention operator check + timing (GPU box).  PGIBBS_ATTN=legacy selects the mma.sync kernel."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from protein_gibbs_sampler_b200.engine import op_attention


def ref(qkv, n_seq, T, H, Dh):
    x = qkv.half().float().cuda().view(n_seq, T, 3, H, Dh)
    q, k, v = (x[:, :, i].transpose(1, 2) for i in range(3))
    return (torch.softmax(q @ k.transpose(-1, -2), -1) @ v).transpose(1, 2).reshape(n_seq * T, H * Dh).cpu()


cases = [(2, 64, 2), (3, 5, 2), (2, 130, 3), (3, 258, 4), (1, 1024, 2), (2, 514, 3), (2, 16, 1), (2, 40, 20), (1, 129, 2)]
for n_seq, T, H in cases:
    qkv = torch.randn(n_seq * T, 3 * H * 64, generator=torch.Generator().manual_seed(T)) * 0.7
    qkv[:, :H * 64] *= 3.0   # sharper softmax
    got = op_attention(qkv, n_seq, T, H, 64)
    want = ref(qkv, n_seq, T, H, 64)
    err = ((got - want).abs().max() / want.abs().max()).item()
    bad = (~torch.isfinite(got)).sum().item()
    print("n_seq %d T %d H %d : rel err %.2e nonfinite %d %s" % (n_seq, T, H, err, bad, "ok" if err < 3e-3 and not bad else "FAIL"), flush=True)
for n_seq, T, H in [(64, 258, 20), (64, 514, 20), (16, 1024, 20)]:
    qkv = torch.randn(n_seq * T, 3 * H * 64, generator=torch.Generator().manual_seed(1)) * 0.7
    got, ms = op_attention(qkv, n_seq, T, H, 64, reps=20)
    want = ref(qkv, n_seq, T, H, 64)
    err = ((got - want).abs().max() / want.abs().max()).item()
    fl = 4.0 * n_seq * H * T * T * 64
    print("n_seq %d T %d H %d : %.3f ms  %.1f TFLOP/s  rel err %.2e" % (n_seq, T, H, ms, fl / ms / 1e9, err), flush=True)
