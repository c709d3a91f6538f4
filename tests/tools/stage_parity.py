"""Stage-by-stage parity of the residual stream (GPU box): engine vs fp32 oracle vs fp16-operand oracle.
    python tests/tools/stage_parity.py [arch] [layers]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from oracle.fair_esm import OracleModel
from fp16_operands import with_fp16_operands
from protein_gibbs_sampler_b200 import models
from protein_gibbs_sampler_b200.config import tiny_config
from protein_gibbs_sampler_b200.esm_sampler import ESM_sampler
from protein_gibbs_sampler_b200.weights import synthetic_state_dict
from test_gpu import _tokens

arch = sys.argv[1] if len(sys.argv) > 1 else "esm2"
layers = int(sys.argv[2]) if len(sys.argv) > 2 else 2
cfg = tiny_config(arch, layers, 1280, 20, 5120)
sd = synthetic_state_dict(cfg, 3)
m = models.CustomModel(cfg, state_dict=sd)
s = ESM_sampler(m, device="cuda:0")
tok = _tokens(cfg, (2, 40), 5)
taps = {}
for name, om in (("fp32", OracleModel(cfg, sd)), ("emul", with_fp16_operands(OracleModel(cfg, sd)))):
    t = {}
    om.model.hook = lambda n, x, t=t: t.__setitem__(n, x.clone())
    t["logits"] = om.model(tok)["logits"]
    taps[name] = t
eng = m.model.engine
M, d = tok.numel(), cfg["embed_dim"]
def rel(a, b): return ((a - b).abs().max() / b.abs().max()).item()
for lim, tap in [(0, "embed")] + [(i + 1, "layer%d" % i) for i in range(layers)]:
    eng.debug_layer_limit(lim)
    eng.forward_logits(tok)
    x = eng.debug_read("x", M * d).view(2, 40, d)
    print("%-8s x: vs fp32 %.3e  vs emul %.3e   (emul vs fp32 %.3e)  max|x| %.3f" % (
        tap, rel(x, taps["fp32"][tap]), rel(x, taps["emul"][tap]), rel(taps["emul"][tap], taps["fp32"][tap]), x.abs().max()))
eng.debug_layer_limit(-1)
lg = eng.forward_logits(tok)
print("logits  : vs fp32 %.3e  vs emul %.3e   (emul vs fp32 %.3e)  max|logit| %.3f" % (
    rel(lg, taps["fp32"]["logits"]), rel(lg, taps["emul"]["logits"]), rel(taps["emul"]["logits"], taps["fp32"]["logits"]), lg.abs().max()))
# per-row error of the logits
e = (lg - taps["emul"]["logits"]).abs().amax(-1) / taps["emul"]["logits"].abs().max()
print("per-token max err vs emul (chain 0):", ["%.1e" % v for v in e[0, :12]])
print("per-token max err vs emul (chain 1):", ["%.1e" % v for v in e[1, :12]])
