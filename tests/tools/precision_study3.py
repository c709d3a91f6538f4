"""CPU emulation of the split-weight ("precise") GEMM mode: which weight classes have to be carried as w_hi + w_lo
(exact to ~2^-22) for the ESM-2 650M logits to stay under 1e-3 of the fp32 oracle, over several weight / token seeds.
Everything else (activations, attention operands, the remaining weights) is rounded to fp16 as the engine does.
    python tests/tools/precision_study3.py [esm2|roberta_large] [T]
"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import torch

from oracle.fair_esm import OracleModel
from protein_gibbs_sampler_b200.config import tiny_config
from protein_gibbs_sampler_b200.weights import synthetic_state_dict
from precision_study import tokens

arch = sys.argv[1] if len(sys.argv) > 1 else "esm2"
T = int(sys.argv[2]) if len(sys.argv) > 2 else 40
cfg = tiny_config(arch, 33, 1280, 20, 5120)
rnd = lambda t: t.half().float()


def cls_of(prefix):
    if prefix.endswith(("q_proj", "k_proj", "v_proj")): return "qkv"
    if prefix.endswith("out_proj"): return "out"
    if prefix.endswith("fc1"): return "fc1"
    if prefix.endswith("fc2"): return "fc2"
    return "head"


def run(sd, tok, ref, exact_w):
    m = OracleModel(cfg, sd)
    mod = m.model

    def lin(x, prefix):
        w = mod.sd[prefix + ".weight"]
        b = w if cls_of(prefix) in exact_w else rnd(w)
        return torch.matmul(rnd(x), b.t()) + mod.sd[prefix + ".bias"]
    mod._lin = lin
    mod.mm = lambda a, b: torch.matmul(rnd(a), rnd(b))
    got = mod(tok)["logits"]
    d = (got - ref).abs()
    per_row = (d.amax(-1) / ref.abs().amax(-1)).max().item()
    return (d.max() / ref.abs().max()).item(), (d.pow(2).mean().sqrt() / ref.abs().max()).item(), per_row


subsets = [(), ("fc2",), ("fc2", "out"), ("fc2", "fc1"), ("fc2", "out", "fc1"), ("fc2", "out", "fc1", "qkv", "head")]
for ws, ts in [(3, 5), (0, 1), (7, 11)]:
    sd = synthetic_state_dict(cfg, ws)
    tok = tokens((2, T), ts)
    ref = OracleModel(cfg, sd).model(tok)["logits"]
    for s in subsets:
        e = run(sd, tok, ref, set(s))
        print("seeds (%d,%d) exact w %-24s: batch-max %.3e rms %.3e per-row-max %.3e" % (ws, ts, "+".join(s) or "-", *e), flush=True)
