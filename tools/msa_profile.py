"""Per-kernel-class time shares of one Gibbs iteration at BASELINE config 3 (MSA-1b, 16 x 32 x 129) (GPU box)."""
import os, random, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from bench import seeds
from protein_gibbs_sampler_b200 import models
from protein_gibbs_sampler_b200.esm_msa_sampler import ESM_MSA_sampler
B, R, L = 16, 32, 128
model = models.ESM_MSA1()
s = ESM_MSA_sampler(model, device="cuda:0")
eng = model.model.engine
rows = seeds(R, L)
toks = model.batch_converter([[(str(i), q) for i, q in enumerate(rows)]] * B)[2]
idx, _ = s.calculate_indexes(None, 0, L, False)
random.seed(0)
plan, _ = s.plan_positions(B, R, idx, -1, 12, False, 6)
eng.set_tokens(toks)
eng.set_schedule(plan.positions, plan.n_iters, plan.P, plan.iter_stride, plan.chain_stride)
eng.set_noise(None); eng.set_device_rng(1)
eng.run(0, 2, float("inf"), 0, None, True, s.valid_aa_idx)
eng.sync()
eng.profile_enable(True)
eng.run(2, 3, float("inf"), 0, None, True, s.valid_aa_idx)
eng.sync()
prof = eng.profile_read()
tot = sum(v[0] for v in prof.values())
for k, (ms, n) in sorted(prof.items(), key=lambda kv: -kv[1][0]):
    print("%-20s %8.3f ms/iter  %5d launches/iter  %7.1f us each  share %.3f" % (k, ms / 3, n // 3, 1000 * ms / n, ms / tot))
print("total %.2f ms/iter" % (tot / 3))
