"""CPU emulation of the engine's operand rounding: which rounded tensor class contributes how much of the logit
error against the fp32 oracle (full-depth models, synthetic weights).  Not part of the product; run by hand:
    python tests/tools/precision_study.py [esm2|roberta_large] [fmt: f16|bf16]
"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch

from oracle.fair_esm import OracleModel
from protein_gibbs_sampler_b200.config import tiny_config
from protein_gibbs_sampler_b200.weights import synthetic_state_dict

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def tokens(shape, seed):
    g = torch.Generator().manual_seed(seed)
    tok = torch.randint(4, 24, shape, generator=g)
    tok[..., 0] = 0
    tok[..., -1] = 2
    flat = tok.view(-1, shape[-1])
    flat[0, 3:9] = 32
    flat[1, 1:shape[-1] - 1] = 32
    return tok


def main():
    arch = sys.argv[1] if len(sys.argv) > 1 else "esm2"
    dt = {"f16": torch.float16, "bf16": torch.bfloat16}[sys.argv[2] if len(sys.argv) > 2 else "f16"]
    cfg = tiny_config(arch, 33, 1280, 20, 5120)
    sd = synthetic_state_dict(cfg, 3)
    tok = tokens((2, 40), 5)
    ref = OracleModel(cfg, sd).model(tok)["logits"]

    def rnd(t):
        return t.to(dt).float()

    def run(w, act, att, split_act=False):
        m = OracleModel(cfg, sd)

        def mm(a, b):
            if b.dim() == 2:  # linear
                bb = rnd(b) if w else b
                if act and split_act:
                    hi = rnd(a)
                    lo = rnd(a - hi)
                    return torch.matmul(hi, bb) + torch.matmul(lo, bb)
                return torch.matmul(rnd(a) if act else a, bb)
            if att:
                return torch.matmul(rnd(a), rnd(b))
            return torch.matmul(a, b)
        m.model.mm = mm
        got = m.model(tok)["logits"]
        per_chain = [((got[i] - ref[i]).abs().max() / ref.abs().max()).item() for i in range(got.shape[0])]
        return ((got - ref).abs().max() / ref.abs().max()).item(), per_chain

    for name, kw in [("weights only", dict(w=1, act=0, att=0)), ("linear inputs only", dict(w=0, act=1, att=0)),
                     ("attention q,k,p,v only", dict(w=0, act=0, att=1)), ("all", dict(w=1, act=1, att=1)),
                     ("all, linear inputs hi+lo", dict(w=1, act=1, att=1, split_act=True))]:
        e, pc = run(**kw)
        print("%-28s max|d|/max|logit| = %.3e   per chain %s" % (name, e, ["%.2e" % v for v in pc]))


if __name__ == "__main__":
    main()
