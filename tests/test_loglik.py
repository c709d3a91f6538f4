"""Pseudo-log-likelihood scoring (shares the forward with the Gibbs step).

Golden values come from the unmodified reference samplers driving the fp32 oracle model
(`tests/golden/make_golden_loglik.py` -> `tests/golden/reference_loglik.json`).
CPU: this package's sampler classes over the SAME oracle model must reproduce them (host logic: strided masking,
batching, gap handling, output order, float32 mean).  GPU: the same classes over the engine, within the forward's
fp16-operand tolerance."""
import json
import os

import pytest
import torch

HERE = os.path.dirname(os.path.abspath(__file__))


@pytest.fixture(scope="module")
def loglik():
    with open(os.path.join(HERE, "golden", "reference_loglik.json")) as f:
        return json.load(f)


def _kw(kw):
    return {k: (float("inf") if v == "inf" else v) for k, v in kw.items()}


def _check(got, want, tol_each, tol_mean):
    assert len(got) == len(want)
    for (gm, gl), (wm, wl) in zip(got, want):
        assert len(gl) == len(wl)
        assert gm == pytest.approx(wm, abs=tol_mean)
        assert gl == pytest.approx(wl, abs=tol_each)


def test_esm_log_likelihood_host_logic_vs_reference(loglik):
    from oracle.fair_esm import OracleModel
    from protein_gibbs_sampler_b200.esm_sampler import ESM_sampler
    from protein_gibbs_sampler_b200.weights import synthetic_state_dict
    for c in loglik["esm"]:
        s = ESM_sampler(OracleModel(c["cfg"], synthetic_state_dict(c["cfg"], c["weights_seed"])), device="cpu")
        _check(list(s.log_likelihood_batch(c["seqs"], **_kw(c["kwargs"]))), c["result"], 2e-6, 2e-6)
    c = loglik["esm"][0]
    s = ESM_sampler(OracleModel(c["cfg"], synthetic_state_dict(c["cfg"], c["weights_seed"])), device="cpu")
    mean, each = s.log_likelihood(c["seqs"][0])
    assert mean == pytest.approx(c["result"][0][0], abs=2e-6) and len(each) == len(c["seqs"][0])
    with pytest.raises(Exception, match="Invalid input character"):
        s.log_likelihood("MKT-AY")


def test_msa_log_likelihood_host_logic_vs_reference(loglik):
    from oracle.fair_esm import OracleModel
    from protein_gibbs_sampler_b200.esm_msa_sampler import ESM_MSA_sampler
    from protein_gibbs_sampler_b200.weights import synthetic_state_dict
    for c in loglik["msa"]:
        s = ESM_MSA_sampler(OracleModel(c["cfg"], synthetic_state_dict(c["cfg"], c["weights_seed"])), device="cpu")
        _check(list(s.log_likelihood_batch(c["msas"], **_kw(c["kwargs"]))), c["result"], 2e-6, 2e-6)
    c = loglik["msa"][0]
    s = ESM_MSA_sampler(OracleModel(c["cfg"], synthetic_state_dict(c["cfg"], c["weights_seed"])), device="cpu")
    mean, each = s.log_likelihood(c["msas"][0], **_kw(c["kwargs"]))
    assert mean == pytest.approx(c["result"][0][0], abs=2e-6)
    assert len(each) == len(c["msas"][0][0].replace("-", ""))   # gaps of the target row are not scored by default


@pytest.mark.gpu
def test_log_likelihood_on_engine_vs_reference(loglik, gpu_lib):
    """Same calls with the CUDA forward: log-probabilities within 3e-3 of the fp32 reference values (a logit error
    of 1e-3 of the largest logit, ~5, moves a log-probability by at most about twice that)."""
    from protein_gibbs_sampler_b200 import models
    from protein_gibbs_sampler_b200.esm_msa_sampler import ESM_MSA_sampler
    from protein_gibbs_sampler_b200.esm_sampler import ESM_sampler
    from protein_gibbs_sampler_b200.weights import synthetic_state_dict
    assert torch.cuda.is_available()
    for key, cls, arg in (("esm", ESM_sampler, "seqs"), ("msa", ESM_MSA_sampler, "msas")):
        cache = {}
        for c in loglik[key]:
            k = json.dumps(c["cfg"], sort_keys=True)
            if k not in cache:
                m = models.CustomModel(c["cfg"], state_dict=synthetic_state_dict(c["cfg"], c["weights_seed"]))
                cache[k] = cls(m, device="cuda:0")
            _check(list(cache[k].log_likelihood_batch(c[arg], **_kw(c["kwargs"]))), c["result"], 3e-3, 1e-3)
