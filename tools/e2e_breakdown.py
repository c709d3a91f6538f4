"""Where does the end-to-end time of one generate() call go?  (GPU box)"""
import os, sys, time, random
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
from protein_gibbs_sampler_b200 import models
from protein_gibbs_sampler_b200.esm_sampler import ESM_sampler

t = time.perf_counter(); m = models.ESM1b(seed=0); print("synthetic weights %.2fs" % (time.perf_counter() - t))
t = time.perf_counter(); s = ESM_sampler(m, device="cuda:0"); print("engine create+load %.2fs" % (time.perf_counter() - t))
seeds = bench.seeds(64, 256)
for K in (3, 10, 10, 40):
    torch.cuda.synchronize(); t0 = time.perf_counter()
    toks = s.get_init_seq(seeds, 256, 64); t1 = time.perf_counter()
    out = s.generate(64, seeds, batch_size=64, num_iters=K, top_k=3, burnin=0, show_progress_bar=False)
    t2 = time.perf_counter()
    print("K=%d total %.3fs (tokenise alone %.3fs) %s -> %.1f it/s" % (K, t2 - t1, t1 - t0, {k: round(v, 4) for k, v in s.last_timing.items()}, K / (t2 - t1)))
