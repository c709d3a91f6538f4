"""Host side of the drop-in boundary (CPU only): tokenisation, index/mask/partition logic, schedule planning
and error behaviour of protein_gibbs_sampler_b200 against the reference's test fixtures
(/root/reference/test/test_esm_sampler.py, test_esm_msa_sampler.py) and the golden vectors."""
import random
import re

import pytest
import torch

from protein_gibbs_sampler_b200 import models
from protein_gibbs_sampler_b200.alphabet import Alphabet, rawbatchlen
from protein_gibbs_sampler_b200.config import tiny_config
from protein_gibbs_sampler_b200.esm_msa_sampler import ESM_MSA_sampler, partition
from protein_gibbs_sampler_b200.esm_sampler import ESM_ALLOWED_AMINO_ACIDS, ESM_sampler


@pytest.fixture(scope="module")
def sampler():
    return ESM_sampler(models.CustomModel(tiny_config("esm2", 1, 64, 2, 128)), device="cpu")


@pytest.fixture(scope="module")
def msa_sampler():
    return ESM_MSA_sampler(models.CustomModel(tiny_config("msa_transformer", 1, 64, 2, 128)), device="cpu")


def test_alphabet_token_ids():
    for a in (Alphabet.esm1b(), Alphabet.msa()):
        assert [a.get_idx(t) for t in ("<cls>", "<pad>", "<eos>", "<unk>", "L", "A", "C", "D", "E", "B", "-", "<mask>")] \
            == [0, 1, 2, 3, 4, 5, 23, 13, 9, 25, 30, 32]
        assert len(a) == 33 and a.get_tok(31) == "<null_1>"
    assert rawbatchlen("AC<mask><mask>-") == 5


def test_valid_ids(sampler, msa_sampler, golden):
    assert sampler.valid_aa_idx == golden["fixtures"]["valid_aa_idx"]["esm"] == list(range(4, 24))
    assert msa_sampler.valid_aa_idx == golden["fixtures"]["valid_aa_idx"]["msa"] == list(range(4, 24)) + [30]
    assert msa_sampler.toks == [sampler.model.alphabet.get_tok(i) for i in msa_sampler.valid_aa_idx]


def test_get_init_seq_golden(sampler, golden):
    for c in golden["fixtures"]["init_seq"]:
        random.seed(c["py_seed"])
        assert sampler.get_init_seq(c["seed_seq"], c["max_len"], c["batch_size"]).tolist() == c["tokens"]
    # reference test_esm_sampler.py:43-60 restated for the ESM-1b alphabet (<cls>=0, <mask>=32, <eos>=2)
    assert sampler.get_init_seq("", 5, 1).tolist() == [[0, 32, 32, 32, 32, 32, 2]]
    assert sampler.get_init_seq("aa", 5, 1).tolist() == [[0, 5, 5, 32, 32, 32, 2]]


def test_get_init_seq_errors(sampler):
    with pytest.raises(Exception) as e:
        sampler.get_init_seq("X", 5, 1)
    assert str(e.value) == "Invalid input character: X"      # test_esm_sampler.py:64-70
    with pytest.raises(Exception) as e:
        sampler.get_init_seq(5, 5, 1)
    assert str(e.value) == "seed sequence should either be a string or list"
    with pytest.raises(ValueError) as e:
        sampler.generate(1, 5)
    assert str(e.value) == "Unknown seed sequence format, expecting str or list"


def test_list_of_seeds_builds_batch_randomly(sampler):
    out = sampler.get_init_seq(["AA", "A"], 5, 3).tolist()   # test_esm_sampler.py:80-88
    assert len(out) == 3
    for row in out:
        assert row in ([0, 5, 5, 32, 32, 32, 2], [0, 5, 32, 32, 32, 32, 2])


def test_device_errors():
    m = models.CustomModel(tiny_config("esm2", 1, 64, 2, 128))
    if not torch.cuda.is_available():
        with pytest.raises(Exception) as e:
            ESM_sampler(m, device="gpu")
        assert str(e.value) == "gpu requested, but No Cuda devices found"
    with pytest.raises(Exception) as e:
        ESM_sampler(m, device="tpu")
    assert str(e.value) == "Invalid device: tpu"


def test_no_cpu_compute_path(sampler):
    """generate() on a CPU-placed sampler must fail loudly, never fall back."""
    with pytest.raises(Exception) as e:
        sampler.generate(1, "MKV", num_iters=1, show_progress_bar=False)
    assert "no CPU path" in str(e.value)


def test_in_order_targets(sampler, golden):
    last_i, t = sampler.get_target_index_in_order(2, [0, 1, 2, 3], 1, 2)   # test_esm_sampler.py:130-137
    assert last_i == 3 and t == [[2, 3], [2, 3]]
    for c in golden["fixtures"]["in_order"]:
        last_i, t = sampler.get_target_index_in_order(2, c["indexes"], c["next_i"], c["num_positions"])
        assert last_i == c["last_i"] and t == c["targets"]


def test_random_targets_follow_python_rng(sampler, golden):
    c = golden["fixtures"]["random_targets"]
    random.seed(c["py_seed"])
    assert sampler.get_random_target_index(c["batch_size"], range(1, 11), c["num_positions"]) == c["targets"]


def test_mask_target_indexes_accepts_nested_lists(sampler, msa_sampler):
    batch = [[1, 2, 3, 4], [5, 6, 7, 8]]                      # test_esm_sampler.py:152-163
    sampler.mask_target_indexes(batch, [[0, 2], [3]])
    assert batch == [[32, 2, 32, 4], [5, 6, 7, 32]]
    mb = [[[1, 2, 3], [4, 5, 6]]]                             # test_esm_msa_sampler.py:168-182
    msa_sampler.mask_target_indexes(mb, [[[0], [1, 2]]])
    assert mb == [[[32, 2, 3], [4, 32, 32]]]
    msa_sampler.mask_target_indexes_single(mb, [1], -1)
    assert mb == [[[32, 2, 3], [4, 32, 32]]]


def test_calculate_indexes(sampler, msa_sampler, golden):
    for c in golden["fixtures"]["calculate_indexes"]:
        out, last = sampler.calculate_indexes(c["indexes"], c["leader"], c["max_len"], c["rollover"])
        assert list(out) == c["out"] and last == c["last_i"]
        out, last = msa_sampler.calculate_indexes(c["indexes"], c["leader"], c["max_len"], c["rollover"])
        assert list(out) == c["out"] and last == c["last_i"]
    # test_esm_msa_sampler.py:185-218
    assert msa_sampler.calculate_indexes(None, 1, 5, False) == ([2, 3, 4, 5], 0)
    assert msa_sampler.calculate_indexes(None, 1, 5, True) == ([1, 2, 3, 4, 5], -1)


def test_partition(golden):
    for c in golden["fixtures"]["partition"]:                 # test_esm_msa_sampler.py:538-557
        assert partition(list(range(c["n"])), c["k"]) == c["out"]
    assert partition([1, 2, 3, 4, 5], 2) == [[1, 2, 3], [4, 5]]


def test_init_msa(msa_sampler, golden):
    c = golden["fixtures"]["init_msa"]
    assert msa_sampler.get_init_msa(c["msa"], c["max_len"], c["batch_size"]).tolist() == c["tokens"]
    with pytest.raises(Exception) as e:
        msa_sampler.get_init_msa(["AX"], 3, 1)
    assert str(e.value) == "Invalid input character: X"
    with pytest.raises(RuntimeError):
        msa_sampler.model.batch_converter([[("0", "AAA"), ("1", "AA")]])


def test_msa_target_layouts(msa_sampler):
    t = msa_sampler.get_target_indexes_all_positions(2, [1, 2, 3], 2)      # test_esm_msa_sampler.py:132-165
    assert t == [[[1, 2, 3], [1, 2, 3]], [[1, 2, 3], [1, 2, 3]]]
    last_i, t = msa_sampler.get_target_index_in_order(2, [0, 1, 2, 3], 1, 2, 3)
    assert last_i == 3 and t == [[[2, 3]] * 3] * 2
    t = msa_sampler.get_random_target_index(2, [1, 2, 3, 4], 2, 3)
    assert len(t) == 2 and all(len(r) == 3 and all(len(p) == 2 and set(p) <= {1, 2, 3, 4} for p in r) for r in t)


class _Recorder:
    """Capture the schedule a sampler would ship to the GPU; pretend no residue changes."""

    def __init__(self):
        self.plans = []

    def __call__(self, tokens, plan, top_k, temperature, burnin, mask):
        self.plans.append(plan)
        return tokens if tokens.dim() == 3 else tokens[:, None, :]


def test_schedule_is_bit_identical_to_reference_masking(golden):
    """Same Python seed -> the pre-drawn schedule equals the positions the reference masked, every iteration."""
    for c in golden["cases"]:
        if not c["kwargs"].get("mask", True):
            continue
        s = ESM_sampler(models.CustomModel(c["cfg"]), device="cpu")
        rec = _Recorder()
        s.run_plan = rec
        random.seed(c["rng_seed"])
        torch.manual_seed(c["rng_seed"])
        s.generate(**c["kwargs"])
        bs, iters = c["kwargs"]["batch_size"], c["kwargs"]["num_iters"]
        got = [[p.targets(it, b) for b in range(bs)] for p in rec.plans for it in range(iters)]
        assert got == c["targets"]
    for c in golden["msa_cases"]:
        if not c["kwargs"].get("mask", True):
            continue
        s = ESM_MSA_sampler(models.CustomModel(c["cfg"]), device="cpu")
        rec = _Recorder()
        s.run_plan = rec
        random.seed(c["rng_seed"])
        torch.manual_seed(c["rng_seed"])
        s.generate(**c["kwargs"])
        bs, iters, R = c["kwargs"]["batch_size"], c["kwargs"]["num_iters"], len(c["kwargs"]["seed_msa"])
        got = [[[p.targets(it, b * R + r) for r in range(R)] for b in range(bs)] for p in rec.plans
               for it in range(iters)]
        assert got == c["targets"]


def test_generate_bookkeeping_without_gpu():
    """Count/truncation/clamping logic of generate (reference :184-207,236-239) with the GPU step stubbed."""
    s = ESM_sampler(models.CustomModel(tiny_config("esm2", 1, 64, 2, 128)), device="cpu")
    s.run_plan = _Recorder()
    out = s.generate(4, "AAAAAAAAAA", batch_size=3, max_len=10, num_iters=2, num_positions=50, leader_length=-1,
                     show_progress_bar=False)
    assert len(out) == 4 and all(len(x) == 10 for x in out)
    assert s.run_plan.plans[0].P == 10                       # clamped to len(indexes)
    out = s.generate(4, "", batch_size=10, max_len=10, num_iters=1, show_progress_bar=False)
    assert len(out) == 4 and out[0] == "<mask>" * 10
    m = ESM_MSA_sampler(models.CustomModel(tiny_config("msa_transformer", 1, 64, 2, 128)), device="cpu")
    m.run_plan = _Recorder()
    out = m.generate(5, ["AC-", "AAA"], batch_size=2, num_iters=1, show_progress_bar=False)
    assert len(out) == 5 and out[:2] == ["AC-", "AAA"]


def test_untokenize(sampler, msa_sampler):
    t = torch.tensor([[0, 5, 23, 32, 2]])
    assert sampler.untokenize_batch(t, True, True) == ["AC<mask>"]
    assert sampler.untokenize_batch(t, True, False) == ["AC<mask><eos>"]
    assert msa_sampler.untokenize_batch(torch.tensor([[[0, 5, 30], [0, 23, 13]]])) == ["A-", "CD"]
    assert set(ESM_ALLOWED_AMINO_ACIDS) == set("ACDEFGHIKLMNPQRSTVWY")


def test_upgrade_state_dict_strips_fair_esm_prefixes():
    sd = {"encoder.sentence_encoder.layers.0.fc1.weight": 1, "encoder.lm_head.bias": 2, "msa.embed_tokens.weight": 3}
    assert set(models.upgrade_state_dict(sd)) == {"layers.0.fc1.weight", "lm_head.bias", "embed_tokens.weight"}


def _as_published(sd, arch):
    """Rename an engine-side (= post-upgrade fair-esm) state dict the way the published checkpoints name their keys:
    the inverse of esm/pretrained.py's key rewriting."""
    out = {}
    for k, v in sd.items():
        if arch == "msa_transformer":
            k = k.replace("row", "column") if "row" in k else k.replace("column", "row")
        if arch == "esm1":
            k = "decoder." + k
        elif arch == "esm2":
            k = ("encoder." if k.startswith("lm_head") else "encoder.sentence_encoder.") + k
        else:
            k = ("encoder." if k.startswith("lm_head") else "encoder.sentence_encoder.") + k
        out[k] = v
    return out


@pytest.mark.parametrize("arch", ["roberta_large", "esm2", "esm1", "msa_transformer"])
def test_upgrade_state_dict_per_architecture(arch, tmp_path):
    """fair-esm's per-architecture key rewriting (esm/pretrained.py, reached by the reference at models.py:61-86): the
    MSA Transformer's published keys have row / column attention swapped, ESM-1 keys carry `decoder.`, ESM-1b zeroes the
    <mask> embedding row.  A checkpoint written with published names must load back to exactly the engine's tensors."""
    import argparse
    import torch
    from protein_gibbs_sampler_b200.config import tiny_config
    from protein_gibbs_sampler_b200.weights import synthetic_state_dict
    cfg = tiny_config(arch, 2, 64, 2, 128)
    sd = {k: v for k, v in synthetic_state_dict(cfg, 1).items()}
    published = _as_published(sd, arch)
    if arch == "msa_transformer":
        # the swap is real: a published "column" key holds what the model calls the tied ROW attention
        assert "encoder.sentence_encoder.layers.0.column_self_attention.layer.q_proj.weight" in published
        assert torch.equal(published["encoder.sentence_encoder.layers.0.column_self_attention.layer.q_proj.weight"],
                           sd["layers.0.row_self_attention.layer.q_proj.weight"])
    if arch == "esm1":
        published["decoder.embed_positions._float_tensor"] = torch.zeros(1)
    up = models.upgrade_state_dict(published, arch)
    assert set(sd) <= set(up)
    for k, v in sd.items():
        if arch == "roberta_large" and k in ("embed_tokens.weight", "lm_head.weight"):
            assert torch.equal(up[k][:32], v[:32]) and not up[k][32].any() and v[32].any()
        else:
            assert torch.equal(up[k], v), k
    # the .pt path: fair-esm checkpoints are {"args": Namespace(arch=...), "model": state dict}
    ck_arch = {"roberta_large": "roberta_large", "esm1": "protein_bert_base", "msa_transformer": "msa_transformer",
               "esm2": None}[arch]
    blob = {"model": published}
    if ck_arch:
        blob["args"] = argparse.Namespace(arch=ck_arch, layers=2)
    path = str(tmp_path / "ck.pt")
    torch.save(blob, path)
    loaded = models.load_checkpoint(path, arch)
    assert all(torch.equal(loaded[k], up[k]) for k in up)
    if ck_arch:
        other = "esm1" if arch != "esm1" else "roberta_large"
        with pytest.raises(Exception, match="expects"):
            models.load_checkpoint(path, other)


def test_checkpoint_loader_refuses_arbitrary_pickles(tmp_path, monkeypatch):
    """A .pt file is user-supplied data: objects beyond tensors / containers / argparse.Namespace are refused unless
    PGIBBS_TRUST_CHECKPOINT=1."""
    import torch

    path = str(tmp_path / "evil.pt")
    torch.save({"model": {"w": torch.zeros(1)}, "hook": _Unpicklable()}, path)
    monkeypatch.delenv("PGIBBS_TRUST_CHECKPOINT", raising=False)
    with pytest.raises(Exception, match="PGIBBS_TRUST_CHECKPOINT"):
        models.load_checkpoint(path, "esm2")
    monkeypatch.setenv("PGIBBS_TRUST_CHECKPOINT", "1")
    assert set(models.load_checkpoint(path, "esm2")) == {"w"}


class _Unpicklable:
    pass


def test_esm1_alphabet_ids_pinned_by_reference_fixtures():
    """ESM-1 alphabet (esm6 / esm12 / esm34): <cls>=32, <mask>=33, A=5 (`/root/reference/test/test_esm_sampler.py:43-66`),
    bos only, 35 tokens; identical to the oracle's restatement of `esm.data.Alphabet.from_architecture("ESM-1")`."""
    from oracle.fair_esm import Alphabet as OracleAlphabet
    from protein_gibbs_sampler_b200.alphabet import Alphabet
    a = Alphabet.esm1()
    assert (len(a), a.cls_idx, a.mask_idx, a.get_idx("A"), a.padding_idx, a.get_idx("<sep>")) == (35, 32, 33, 5, 1, 34)
    assert a.prepend_bos and not a.append_eos
    assert a.all_toks == OracleAlphabet.from_architecture("ESM-1").all_toks
    toks = a.get_batch_converter()([("0", "AA<mask>")])[2]
    assert toks.tolist() == [[32, 5, 5, 33]]
    assert sorted(a.get_idx(t) for t in "ACDEFGHIKLMNPQRSTVWY") == list(range(4, 24))


def test_strided_mask_plan_covers_every_position_once():
    """Scoring schedule (log_likelihood_batch's `toks[i, start+i : start+L : n] = mask`, reference
    esm_sampler.py:323-326): copy i masks i, i+n, ...; padded slots repeat the copy's first position."""
    import numpy as np
    from protein_gibbs_sampler_b200.esm_sampler import strided_mask_plan
    for L, n, start in [(10, 10, 1), (10, 3, 1), (7, 1, 0), (257, 16, 1), (5, 4, 1)]:
        pos, valid = strided_mask_plan(L, n, start)
        assert pos.shape == valid.shape == (n, -(-L // n))
        assert sorted(pos[valid].tolist()) == list(range(start, start + L))
        for i in range(n):
            assert pos[i][valid[i]].tolist() == list(range(start + i, start + L, n))
            assert (pos[i][~valid[i]] == start + i).all()
