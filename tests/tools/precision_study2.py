"""Finer CPU emulation: contribution of each GEMM class (qkv / out / fc1 / fc2 / head) and operand (W / activation)
to the logit error vs the fp32 oracle, by rounding everything EXCEPT that class-operand (leave-one-out)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch
from oracle.fair_esm import OracleModel
from protein_gibbs_sampler_b200.config import tiny_config
from protein_gibbs_sampler_b200.weights import synthetic_state_dict
from precision_study import tokens

arch = sys.argv[1] if len(sys.argv) > 1 else "esm2"
cfg = tiny_config(arch, 33, 1280, 20, 5120)
sd = synthetic_state_dict(cfg, 3)
tok = tokens((2, 40), 5)
ref = OracleModel(cfg, sd).model(tok)["logits"]
rnd = lambda t: t.half().float()


def cls_of(prefix):
    if prefix.endswith(("q_proj", "k_proj", "v_proj")): return "qkv"
    if prefix.endswith("out_proj"): return "out"
    if prefix.endswith("fc1"): return "fc1"
    if prefix.endswith("fc2"): return "fc2"
    return "head"


def run(exact):  # set of (class, operand) kept exact; "att" for attention matmuls
    m = OracleModel(cfg, sd)
    mod = m.model

    def lin(x, prefix):
        c = cls_of(prefix)
        w = mod.sd[prefix + ".weight"]
        a = x if (c, "act") in exact else rnd(x)
        b = w if (c, "w") in exact else rnd(w)
        return torch.matmul(a, b.t()) + mod.sd[prefix + ".bias"]
    mod._lin = lin
    mod.mm = (lambda a, b: torch.matmul(a, b)) if "att" in exact else (lambda a, b: torch.matmul(rnd(a), rnd(b)))
    got = mod(tok)["logits"]
    d = (got - ref).abs()
    return (d.max() / ref.abs().max()).item(), (d.pow(2).mean().sqrt() / ref.abs().max()).item()

base = run(set())
print("all rounded: max %.3e rms %.3e" % base)
for c in ("qkv", "out", "fc1", "fc2", "head"):
    for o in ("w", "act"):
        e = run({(c, o)})
        print("exact %-4s %-3s : max %.3e rms %.3e   (rms^2 share %.2f)" % (c, o, e[0], e[1], 1 - (e[1] / base[1]) ** 2))
e = run({"att"})
print("exact attention: max %.3e rms %.3e   (rms^2 share %.2f)" % (e[0], e[1], 1 - (e[1] / base[1]) ** 2))
for name, ex in [("fc2 w+act", {("fc2", "w"), ("fc2", "act")}), ("fc1+fc2 w", {("fc1", "w"), ("fc2", "w")}),
                 ("out+fc2 w+act", {("fc2", "w"), ("fc2", "act"), ("out", "w"), ("out", "act")}),
                 ("all w", {(c, "w") for c in ("qkv", "out", "fc1", "fc2", "head")}),
                 ("all act", {(c, "act") for c in ("qkv", "out", "fc1", "fc2", "head")})]:
    e = run(ex)
    print("exact %-14s: max %.3e rms %.3e" % (name, e[0], e[1]))
