"""Attention kernel timing only (GPU box; A/B tooling): python tools/attn_time.py  [PGIBBS_LIB_PATH selects the build]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from protein_gibbs_sampler_b200.engine import op_attention
out = []
for n_seq, T, H in [(64, 258, 20), (64, 514, 20), (16, 1024, 20)]:
    qkv = torch.randn(n_seq * T, 3 * H * 64, generator=torch.Generator().manual_seed(1)) * 0.7
    _, ms = op_attention(qkv, n_seq, T, H, 64, reps=50)
    out.append("T%d %.1f us %.0f TF" % (T, ms * 1000, 4.0 * n_seq * H * T * T * 64 / ms / 1e9))
print(os.path.basename(os.environ.get("PGIBBS_LIB_PATH", "default")), " | ".join(out), flush=True)
