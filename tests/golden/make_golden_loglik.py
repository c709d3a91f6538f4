"""Golden pseudo-log-likelihood values from the UNMODIFIED reference samplers (imported from /root/reference/src)
driving the fp32 CPU oracle model with seeded synthetic weights.  Writes tests/golden/reference_loglik.json.

    python tests/golden/make_golden_loglik.py        # needs /root/reference (this container only)

The reference's own known-answer tests for this function (`/root/reference/test/test_esm_sampler.py:269-340`,
`test_esm_msa_sampler.py:265-397`) need pretrained weights, which do not exist offline; these vectors pin the same code
paths (strided masking, batching, gap handling, output order) on the oracle model instead.
"""
import json
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, "/root/reference/src")

import pgen.esm_msa_sampler as ref_msa  # noqa: E402
import pgen.esm_sampler as ref_esm  # noqa: E402
from oracle.fair_esm import OracleModel  # noqa: E402
from protein_gibbs_sampler_b200.config import tiny_config  # noqa: E402
from protein_gibbs_sampler_b200.weights import synthetic_state_dict  # noqa: E402


def inf(v):
    return "inf" if v == float("inf") else v


def main():
    out = {"esm": [], "msa": []}
    for arch in ("esm2", "roberta_large"):
        cfg = tiny_config(arch, layers=2, embed_dim=128, heads=2, ffn_dim=256)
        s = ref_esm.ESM_sampler(OracleModel(cfg, synthetic_state_dict(cfg, 7)), device="cpu")
        for seqs, kw in [
            (["MKTAYIAKQRQISFVKSHFSRQ"], dict(with_masking=True)),
            (["MKTAYIAKQRQISFVKSHFSRQ"], dict(with_masking=True, mask_distance=3)),
            (["MKTAYIAKQRQISFVKSHFSRQ"], dict(with_masking=True, mask_distance=5, batch_size=2)),
            (["MKTAYIAKQRQISFVKSHFSRQ"], dict(with_masking=False)),
            (["MKTAYIAKQR", "acdefghikl"], dict(with_masking=True, mask_distance=4)),
            (["MKTAYIAKQR", "ACDEFGHIKL"], dict(with_masking=False)),
        ]:
            res = list(s.log_likelihood_batch(seqs, **kw))
            out["esm"].append({"arch": arch, "cfg": cfg, "weights_seed": 7, "seqs": seqs,
                               "kwargs": {k: inf(v) for k, v in kw.items()}, "result": res})
    cfg = tiny_config("msa_transformer", layers=2, embed_dim=128, heads=2, ffn_dim=256)
    s = ref_msa.ESM_MSA_sampler(OracleModel(cfg, synthetic_state_dict(cfg, 3)), device="cpu")
    msa = ["MKTAYIAK-RQ", "MKSAY-AKQRQ", "MRTAYIAKQ-Q", "M-TAYLAKQRQ"]
    msa2 = ["ACDEFGHIKLM", "ACDEF-HIKLM", "AC-EFGHIKLM", "ACDEFGHIK-M"]
    for msas, kw in [
        ([msa], dict(target_index=0, with_masking=True)),
        ([msa], dict(target_index=0, with_masking=True, count_gaps=True)),
        ([msa], dict(target_index=2, with_masking=True, mask_distance=3, batch_size=2)),
        ([msa], dict(target_index=1, with_masking=True, mask_distance=4, count_gaps=True, batch_size=None)),
        ([msa, msa2], dict(target_index=0, with_masking=False)),
        ([msa, msa2], dict(target_index=3, with_masking=False, count_gaps=True, batch_size=2)),
        ([msa, msa2], dict(target_index=1, with_masking=True, mask_distance=2)),
    ]:
        res = list(s.log_likelihood_batch(msas, **kw))
        out["msa"].append({"cfg": cfg, "weights_seed": 3, "msas": msas, "kwargs": {k: inf(v) for k, v in kw.items()},
                           "result": res})
    path = os.path.join(HERE, "reference_loglik.json")
    with open(path, "w") as f:
        json.dump(out, f)
    print("wrote", path, os.path.getsize(path), "bytes;", len(out["esm"]), "esm cases,", len(out["msa"]), "msa cases")


if __name__ == "__main__":
    main()
