"""CPU emulation of the split-operand GEMM modes with the CORRECTION passes in 8-bit floats:
    y = a_hi w_hi  +  q8(a_hi) q8(w_lo) [weights' rounding error]  (+ q8(a_lo) q8(w_hi) [activations'])
where a_hi = fp16(a), a_lo = a - a_hi, w likewise, and q8 rounds to e4m3 (lo parts, scaled by a power of two per tensor so
that they sit in the normal range) or e5m2 (hi parts: two mantissa bits are enough for a term that is 2^-12 of the sum).
tcgen05 `kind::f8f6f4` runs such a pass at twice the fp16 rate.  Question answered here: do 3-bit corrections recover
what the fp16 hi + lo pairs of the shipped split modes recover?  Everything else is rounded to fp16 as the engine does.
    python tests/tools/precision_study4.py [esm2|roberta_large] [T]
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


def q8(t, dtype):
    """Round to an 8-bit float after a per-tensor power-of-two scale that puts max|t| near 2^7; returns the de-scaled value
    (the scale is exact, so only the 8-bit rounding remains)."""
    m = t.abs().max().item()
    if m == 0.0:
        return t
    s = 2.0 ** (7 - torch.tensor(m).log2().ceil().item())
    return (t * s).to(dtype).float() / s


def run(sd, tok, ref, mode):
    mod = OracleModel(cfg, sd).model

    def lin(x, prefix):
        w = mod.sd[prefix + ".weight"]
        a_hi, w_hi = rnd(x), rnd(w)
        y = torch.matmul(a_hi, w_hi.t())
        if mode in ("w16", "aw16"):      # the shipped split modes: fp16 lo halves
            y = y + torch.matmul(a_hi, rnd(w - w_hi).t())
        if mode == "aw16":
            y = y + torch.matmul(rnd(x - a_hi), w_hi.t())
        if mode in ("w8", "aw8"):        # 8-bit corrections
            y = y + torch.matmul(q8(a_hi, torch.float8_e5m2), q8(w - w_hi, torch.float8_e4m3fn).t())
        if mode == "aw8":
            y = y + torch.matmul(q8(x - a_hi, torch.float8_e4m3fn), q8(w_hi, torch.float8_e5m2).t())
        return y + mod.sd[prefix + ".bias"]
    mod._lin = lin
    mod.mm = lambda a, b: torch.matmul(rnd(a), rnd(b))
    got = mod(tok)["logits"]
    d = (got - ref).abs()
    return ((d.max() / ref.abs().max()).item(), (d.pow(2).mean().sqrt() / ref.abs().max()).item(),
            (d.amax(-1) / ref.abs().amax(-1)).max().item())


NAMES = {"fast": "fast (one fp16 pass)", "w16": "split_weights (fp16 w_lo)", "w8": "weights' correction in fp8",
         "aw16": "split (fp16 a_lo, w_lo)", "aw8": "both corrections in fp8"}
for ws, ts in [(3, 5), (0, 1), (7, 11)]:
    sd = synthetic_state_dict(cfg, ws)
    tok = tokens((2, T), ts)
    ref = OracleModel(cfg, sd).model(tok)["logits"]
    for mode in ("fast", "w16", "w8", "aw16", "aw8"):
        e = run(sd, tok, ref, mode)
        print("seeds (%d,%d) %-28s: batch-max %.3e rms %.3e per-row-max %.3e" % (ws, ts, NAMES[mode], *e), flush=True)
