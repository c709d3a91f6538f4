import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from protein_gibbs_sampler_b200.engine import op_attention
n_seq, T, H = 64, int(os.environ.get("ATT_T", 258)), 20
qkv = torch.randn(n_seq * T, 3 * H * 64, generator=torch.Generator().manual_seed(1)) * 0.7
op_attention(qkv, n_seq, T, H, 64, reps=2)
