"""Device-resident Gibbs-iteration throughput of the per-GPU shard of every BASELINE config (GPU box):
    python tools/config_bench.py [c2 c3 c4 c5] > gpurun_out/config_bench.txt
One line per config: iters/s, ms/iter, algorithmic TFLOP/s and fraction of the measured sustained bf16 peak.
bench.py stays the contract line (config 2); this is the table in BASELINE.md section 5."""
import json, os, random, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from bench import algorithmic_flops_per_iter, measured_peaks, seeds
from protein_gibbs_sampler_b200 import models
from protein_gibbs_sampler_b200.esm_msa_sampler import ESM_MSA_sampler
from protein_gibbs_sampler_b200.esm_sampler import ESM_sampler

PEAK = measured_peaks()["tensor"]


def msa_flops(cfg, B, R, C):
    d, F, V, L = cfg["embed_dim"], cfg["ffn_dim"], cfg["vocab"], cfg["layers"]
    return B * R * C * (L * (2 * (8 * d * d + 2 * d * F) + 4 * C * d + 4 * R * d) + 2 * (d * d + d * V))


def timed(engine, run, warm, iters):
    stream = torch.cuda.Stream()
    torch.cuda.set_stream(stream)
    engine.set_stream(stream.cuda_stream)
    run(0, warm)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    run(warm, iters)
    e1.record(stream)
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters


def single(name, model, B, L, top_k, burnin, num_positions, iters=10, warm=3):
    s = ESM_sampler(model, device="cuda:0")
    eng = model.model.engine
    toks = model.batch_converter([(str(i), q) for i, q in enumerate(seeds(B, L))])[2]
    idx, _ = s.calculate_indexes(None, 0, L, False)
    random.seed(0)
    plan, _ = s.plan_positions(B, idx, -1, num_positions, False, warm + iters)
    eng.set_tokens(toks)
    eng.set_schedule(plan.positions, plan.n_iters, plan.P, plan.iter_stride, plan.chain_stride)
    eng.set_noise(None)
    eng.set_device_rng(1)
    ms = timed(eng, lambda a, n: eng.run(a, n, burnin, top_k, None, True, s.valid_aa_idx), warm, iters)
    fl = algorithmic_flops_per_iter(model.cfg, B, toks.shape[1])
    report(name, ms, fl, "P=%d" % plan.P)
    eng.close()


def msa(name, B, R, L, num_positions, iters=6, warm=2):
    model = models.ESM_MSA1()
    s = ESM_MSA_sampler(model, device="cuda:0")
    eng = model.model.engine
    rows = seeds(R, L)
    toks = model.batch_converter([[(str(i), q) for i, q in enumerate(rows)]] * B)[2]
    idx, _ = s.calculate_indexes(None, 0, L, False)
    random.seed(0)
    plan, _ = s.plan_positions(B, R, idx, -1, num_positions, False, warm + iters)
    eng.set_tokens(toks)
    eng.set_schedule(plan.positions, plan.n_iters, plan.P, plan.iter_stride, plan.chain_stride)
    eng.set_noise(None)
    eng.set_device_rng(1)
    ms = timed(eng, lambda a, n: eng.run(a, n, float("inf"), 0, None, True, s.valid_aa_idx), warm, iters)
    report(name, ms, msa_flops(model.cfg, B, R, toks.shape[2]), "P=%d per row" % plan.P)
    eng.close()


def report(name, ms, flops, note):
    tf = flops / (ms * 1e-3) / 1e12
    print("%-44s %8.2f ms/iter %8.2f iters/s %8.1f algorithmic TFLOP/s  %.3f of sustained peak  (%s)"
          % (name, ms, 1000.0 / ms, tf, tf / PEAK, note), flush=True)


want = set(sys.argv[1:]) or {"c2", "c3", "c4", "c5"}
if "c1" in want:   # launch-bound shapes: BASELINE config 1 geometry and single-chain runs (pgen_msa_revised's regime)
    single("C1 esm2_t6_8M, 2 x L25 (on the GPU)", models.ESM2_t6_8M(), 2, 25, 0, float("inf"), 2, iters=50, warm=5)
    single("   ESM-1b 650M, 1 x L256", models.ESM1b(), 1, 256, 0, float("inf"), 25, iters=20, warm=3)
    msa("   MSA-1b, 1 MSA x 32 rows x L128", 1, 32, 128, 12, iters=20, warm=3)
if "c2" in want:
    single("C2 ESM-1b 650M, 64 x L256, top_k 3", models.ESM1b(), 64, 256, 3, 0, 0)
if "c3" in want:
    msa("C3 MSA-1b, 16 MSAs x 32 rows x L128, 10 %", 16, 32, 128, 12)
    msa("C3 MSA-1b, 16 MSAs x 32 rows x L128, all", 16, 32, 128, 0)
if "c4" in want:
    single("C4 ESM-2 650M shard, 64 x L512, top_k 5", models.ESM2_t33_650M(), 64, 512, 5, 50, 0, iters=6)
if "c5" in want:
    for pct in (5, 10, 25):
        single("C5 ESM-1b shard, 16 x L1022, %d %% positions" % pct, models.ESM1b(), 16, 1022, 0, float("inf"),
               int(1022 * pct / 100), iters=8)
