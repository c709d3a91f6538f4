"""Generate golden vectors by running the UNMODIFIED reference (imported from /root/reference/src) on the
oracle model.  Run in the build container only (the GPU box has no /root/reference):

    python tests/golden/make_golden.py

Writes tests/golden/reference_golden.json.  Everything in it is produced by reference code:
pgen.esm_sampler / pgen.esm_msa_sampler drive oracle.fair_esm models (fair-esm itself is not installable),
with Python's and torch's global RNGs seeded as recorded.
"""
import json
import os
import random
import sys

import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, "/root/reference/src")

from pgen import esm_msa_sampler as ref_msa  # noqa: E402
from pgen import esm_sampler as ref_esm  # noqa: E402

from oracle.fair_esm import OracleModel  # noqa: E402
from protein_gibbs_sampler_b200.config import tiny_config  # noqa: E402
from protein_gibbs_sampler_b200.weights import synthetic_state_dict  # noqa: E402


def traced_generate(sampler, kwargs, seed, msa=False):
    """Run reference generate; record every iteration's target indexes and the token state after it."""
    trace = {"targets": [], "states": [], "init": None}
    model = sampler.model.model
    orig_forward = model.forward

    def fwd(tokens, **kw):
        trace["states"].append(tokens.clone().tolist())  # state fed to the forward (after masking)
        return orig_forward(tokens, **kw)

    model.forward = fwd
    orig_mask = sampler.mask_target_indexes

    def mask_hook(batch, target_indexes):
        trace["targets"].append([list(map(int, t)) if not msa else [list(map(int, r)) for r in t]
                                 for t in target_indexes])
        return orig_mask(batch, target_indexes)

    sampler.mask_target_indexes = mask_hook
    random.seed(seed)
    torch.manual_seed(seed)
    out = sampler.generate(**kwargs)
    model.forward = orig_forward
    sampler.mask_target_indexes = orig_mask
    return out, trace


def main():
    golden = {"cases": [], "msa_cases": [], "single_cases": [], "generate_step": [], "fixtures": {}}

    # ---------------- single-sequence sampler on a tiny ESM-2 and a tiny ESM-1b
    for arch, seedval in (("esm2", 0), ("roberta_large", 1)):
        cfg = tiny_config(arch, layers=2, embed_dim=128, heads=2, ffn_dim=256)
        sd_seed = 7
        om = OracleModel(cfg, synthetic_state_dict(cfg, sd_seed))
        sampler = ref_esm.ESM_sampler(om, device="cpu")
        cases = [
            dict(n_samples=3, seed_seq="MKTAYIAKQRQISFVKSHFSRQ", batch_size=3, num_iters=3, top_k=3, burnin=1,
                 num_positions=5, show_progress_bar=False),
            dict(n_samples=2, seed_seq="MKTAYIAKQR", batch_size=2, max_len=16, num_iters=2, top_k=0,
                 in_order=True, num_positions=3, leader_length=2, show_progress_bar=False),
            dict(n_samples=4, seed_seq=["MKTAYIAK", "MKV", "ACDEFGHIKL"], batch_size=2, max_len=12, num_iters=2,
                 temperature=0.8, top_k=2, burnin=0, mask=False, show_progress_bar=False),
            dict(n_samples=2, seed_seq="MKTAYIAKQRQISFVK", batch_size=2, num_iters=2, top_k=1, burnin=0,
                 num_positions_percent=25, leader_length_percent=10, rollover_from_start=True, in_order=True,
                 show_progress_bar=False),
        ]
        for i, kw in enumerate(cases):
            out, trace = traced_generate(sampler, kw, seed=100 + i)
            golden["cases"].append({"arch": arch, "cfg": cfg, "weights_seed": sd_seed, "rng_seed": 100 + i,
                                    "kwargs": kw, "output": out, "targets": trace["targets"],
                                    "states": trace["states"]})

    # ---------------- MSA sampler on a tiny MSA transformer
    cfg = tiny_config("msa_transformer", layers=2, embed_dim=128, heads=2, ffn_dim=256)
    om = OracleModel(cfg, synthetic_state_dict(cfg, 3))
    msampler = ref_msa.ESM_MSA_sampler(om, device="cpu")
    msa = ["MKTAYIAK-RQ", "MKSAY-AKQRQ", "MRTAYIAKQ-Q", "M-TAYLAKQRQ"]
    mcases = [
        dict(n_samples=8, seed_msa=msa, batch_size=2, num_iters=2, top_k=3, burnin=1, num_positions=3,
             show_progress_bar=False),
        dict(n_samples=4, seed_msa=msa, batch_size=1, num_iters=2, in_order=True, num_positions=2, leader_length=1,
             show_progress_bar=False),
        dict(n_samples=4, seed_msa=msa, batch_size=1, num_iters=1, top_k=1, burnin=0, mask=False,
             show_progress_bar=False),
    ]
    for i, kw in enumerate(mcases):
        out, trace = traced_generate(msampler, kw, seed=200 + i, msa=True)
        golden["msa_cases"].append({"cfg": cfg, "weights_seed": 3, "rng_seed": 200 + i, "kwargs": kw,
                                    "output": out, "targets": trace["targets"], "states": trace["states"]})
    for i, kw in enumerate([dict(seed_msa=msa, steps=3, passes=2, burn_in=1, target_index=0, k=1),
                            dict(seed_msa=msa, steps=2, passes=2, burn_in=0, target_index=-1, k=2,
                                 exclude_positions=[0, 4])]):
        random.seed(300 + i)
        torch.manual_seed(300 + i)
        out = msampler.generate_single(**kw)
        golden["single_cases"].append({"cfg": cfg, "weights_seed": 3, "rng_seed": 300 + i, "kwargs": kw,
                                       "output": out})

    # ---------------- generate_step draws under seeded torch RNG
    g = torch.Generator().manual_seed(5)
    for i in range(40):
        valid = [list(range(4, 24)), [3, 5, 1], list(range(4, 24)) + [30], None][i % 4]
        top_k = [0, 2, 3, 40][i % 4] if valid is not None else [0, 5][i % 2]
        temp = [None, 0.5, 2.0][i % 3]
        sample = i % 5 == 0
        logits = (torch.randn(4, 33, generator=g) * 2).tolist()
        torch.manual_seed(1000 + i)
        idx = int(ref_esm.generate_step(torch.tensor(logits), 2, temperature=temp, top_k=top_k, sample=sample,
                                        valid_idx=valid))
        golden["generate_step"].append({"logits": logits, "gen_idx": 2, "temperature": temp, "top_k": top_k,
                                        "sample": sample, "valid_idx": valid, "torch_seed": 1000 + i, "token": idx})

    # ---------------- exact host-logic fixtures straight from the reference objects
    fx = golden["fixtures"]
    fx["partition"] = [{"n": n, "k": k, "out": ref_msa.partition(list(range(n)), k)}
                       for n in (0, 1, 5, 10, 11) for k in (1, 2, 3, 4, 7, 12) if n > 0]
    cfg1 = tiny_config("esm2", layers=1, embed_dim=64, heads=2, ffn_dim=128)
    s1 = ref_esm.ESM_sampler(OracleModel(cfg1, synthetic_state_dict(cfg1, 0)), device="cpu")
    fx["calculate_indexes"] = []
    for idx, lead, ml, roll in [(None, 1, 5, False), (None, 1, 5, True), ([2, 3, 4, 5], 1, 5, False), (None, 0, 4, False),
                                (None, 3, 8, False)]:
        out, last = s1.calculate_indexes(idx, lead, ml, roll)
        fx["calculate_indexes"].append({"indexes": idx, "leader": lead, "max_len": ml, "rollover": roll,
                                        "out": list(out), "last_i": last})
    fx["in_order"] = []
    for nxt, npos, idx in [(1, 2, [0, 1, 2, 3]), (-1, 3, [1, 2, 3, 4, 5]), (4, 7, [1, 2, 3, 4, 5])]:
        last, t = s1.get_target_index_in_order(2, idx, nxt, npos)
        fx["in_order"].append({"next_i": nxt, "num_positions": npos, "indexes": idx, "last_i": last, "targets": t})
    fx["init_seq"] = []
    for seed_seq, ml, bs, rs in [("", 5, 1, 0), ("AA", 5, 1, 0), ("aa", 5, 2, 0), (["AA", "A", "MKV"], 6, 4, 11)]:
        random.seed(rs)
        fx["init_seq"].append({"seed_seq": seed_seq, "max_len": ml, "batch_size": bs, "py_seed": rs,
                               "tokens": s1.get_init_seq(seed_seq, ml, bs).tolist()})
    random.seed(21)
    fx["random_targets"] = {"py_seed": 21, "batch_size": 3, "indexes": list(range(1, 11)), "num_positions": 4,
                            "targets": s1.get_random_target_index(3, range(1, 11), 4)}
    fx["init_msa"] = {"msa": ["AC-", "a"], "max_len": 4, "batch_size": 2,
                      "tokens": msampler.get_init_msa(["AC-", "a"], 4, 2).tolist()}
    fx["valid_aa_idx"] = {"esm": s1.valid_aa_idx, "msa": msampler.valid_aa_idx}

    with open(os.path.join(HERE, "reference_golden.json"), "w") as f:
        json.dump(golden, f)
    print("wrote", os.path.join(HERE, "reference_golden.json"), os.path.getsize(os.path.join(HERE, "reference_golden.json")), "bytes")


if __name__ == "__main__":
    main()
