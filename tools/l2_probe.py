"""How the per-row cost of the memory-bound kernels depends on the working-set size (GPU box): ESM-1b 650M, L=256,
B chains, per-kernel time from the engine's event profiler divided by the number of token rows.
    python tools/l2_probe.py > gpurun_out/l2_probe.txt
x (fp32 residual) is 1.32 MB per chain, h / ctx 0.66 MB, qkv 1.98 MB, ffn 2.64 MB."""
import os, random, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from bench import seeds
from protein_gibbs_sampler_b200 import models
from protein_gibbs_sampler_b200.esm_sampler import ESM_sampler

model = models.ESM1b()
s = ESM_sampler(model, device="cuda:0")
eng = model.model.engine
L = 256
print("%5s %8s | us per launch (ns per row)" % ("B", "x MB"))
for B in [int(a) for a in sys.argv[1:]] or [4, 8, 16, 24, 32, 48, 64, 96, 128]:
    toks = model.batch_converter([(str(i), q) for i, q in enumerate(seeds(B, L))])[2]
    idx, _ = s.calculate_indexes(None, 0, L, False)
    random.seed(0)
    plan, _ = s.plan_positions(B, idx, -1, 0, False, 8)
    eng.set_tokens(toks)
    eng.set_schedule(plan.positions, plan.n_iters, plan.P, plan.iter_stride, plan.chain_stride)
    eng.set_noise(None)
    eng.set_device_rng(1)
    eng.run(0, 3, 0, 3, None, True, s.valid_aa_idx)
    eng.sync()
    eng.profile_enable(True)
    eng.run(3, 5, 0, 3, None, True, s.valid_aa_idx)
    eng.sync()
    prof = eng.profile_read()
    eng.profile_enable(False)
    rows = B * (L + 2)
    cells = []
    for k in ("layernorm", "attention", "gemm_qkv", "gemm_out", "gemm_fc1", "gemm_fc2"):
        ms, n = prof[k]
        us = 1e3 * ms / n
        cells.append("%s %6.1f (%5.2f)" % (k.replace("gemm_", ""), us, 1e3 * us / rows))
    print("%5d %8.1f | %s" % (B, rows * 1280 * 4 / 1e6, "  ".join(cells)), flush=True)
