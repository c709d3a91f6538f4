"""The oracle against the reference's own fixtures and golden vectors (CPU only).

Sources of truth: token ids and index fixtures asserted in the reference's tests
(/root/reference/test/test_esm_sampler.py:43-88,130-163; test_esm_msa_sampler.py:43-84,132-218,538-557) and
tests/golden/reference_golden.json, produced by tests/golden/make_golden.py running the unmodified reference.
"""
import random

import pytest
import torch

from oracle import gibbs_loop
from oracle.fair_esm import Alphabet, OracleModel
from oracle.sampler_tail import effective_k, generate_step_with_noise
from protein_gibbs_sampler_b200.config import tiny_config
from protein_gibbs_sampler_b200.weights import synthetic_state_dict


def test_alphabet_ids_pinned_by_reference_tests():
    esm1 = Alphabet.from_architecture("ESM-1")           # test_esm_sampler.py:46,53
    assert (esm1.cls_idx, esm1.mask_idx, esm1.get_idx("A")) == (32, 33, 5)
    for arch in ("ESM-1b", "MSA Transformer"):            # test_esm_msa_sampler.py:45-66
        a = Alphabet.from_architecture(arch)
        assert [a.get_idx(t) for t in ("<cls>", "A", "C", "D", "E", "B", "<mask>")] == [0, 5, 23, 13, 9, 25, 32]
        assert len(a) == 33 and a.padding_idx == 1 and a.eos_idx == 2
    assert Alphabet.from_architecture("ESM-1b").append_eos and not Alphabet.from_architecture("MSA Transformer").append_eos


def test_batch_converter_mask_literal_is_one_token():
    a = Alphabet.from_architecture("ESM-1b")
    toks = a.get_batch_converter()([("0", "AA<mask><mask>"), ("1", "A")])[2]
    assert toks.tolist() == [[0, 5, 5, 32, 32, 2], [0, 5, 2, 1, 1, 1]]


def _hf_model(cfg, sd, rotary):
    from transformers import EsmConfig, EsmForMaskedLM
    hc = EsmConfig(vocab_size=33, mask_token_id=32, pad_token_id=1, hidden_size=cfg["embed_dim"],
                   num_hidden_layers=cfg["layers"], num_attention_heads=cfg["heads"],
                   intermediate_size=cfg["ffn_dim"], layer_norm_eps=1e-5, max_position_embeddings=1026,
                   token_dropout=True, position_embedding_type="rotary" if rotary else "absolute",
                   emb_layer_norm_before=not rotary, hidden_dropout_prob=0.0, attention_probs_dropout_prob=0.0)
    m = EsmForMaskedLM(hc).eval()
    hs = m.state_dict()
    mp = {"esm.embeddings.word_embeddings.weight": "embed_tokens.weight",
          "esm.encoder.emb_layer_norm_after.weight": "emb_layer_norm_after.weight",
          "esm.encoder.emb_layer_norm_after.bias": "emb_layer_norm_after.bias",
          "lm_head.dense.weight": "lm_head.dense.weight", "lm_head.dense.bias": "lm_head.dense.bias",
          "lm_head.layer_norm.weight": "lm_head.layer_norm.weight", "lm_head.layer_norm.bias": "lm_head.layer_norm.bias",
          "lm_head.decoder.weight": "embed_tokens.weight", "lm_head.bias": "lm_head.bias",
          "lm_head.decoder.bias": "lm_head.bias"}
    if not rotary:
        mp["esm.embeddings.position_embeddings.weight"] = "embed_positions.weight"
        mp["esm.embeddings.layer_norm.weight"] = "emb_layer_norm_before.weight"
        mp["esm.embeddings.layer_norm.bias"] = "emb_layer_norm_before.bias"
    for i in range(cfg["layers"]):
        h, o = "esm.encoder.layer.%d." % i, "layers.%d." % i
        for a, b in (("attention.self.query", "self_attn.q_proj"), ("attention.self.key", "self_attn.k_proj"),
                     ("attention.self.value", "self_attn.v_proj"), ("attention.output.dense", "self_attn.out_proj"),
                     ("attention.LayerNorm", "self_attn_layer_norm"), ("intermediate.dense", "fc1"),
                     ("output.dense", "fc2"), ("LayerNorm", "final_layer_norm")):
            mp[h + a + ".weight"] = o + b + ".weight"
            mp[h + a + ".bias"] = o + b + ".bias"
    new = {}
    for k, v in hs.items():
        if k in mp:
            new[k] = sd[mp[k]].clone()
        else:
            assert "inv_freq" in k or "position_ids" in k or "contact_head" in k, k
            new[k] = v
    m.load_state_dict(new)
    return m


@pytest.mark.parametrize("arch", ["esm2", "roberta_large"])
def test_forward_matches_independent_hf_implementation(arch):
    """transformers' EsmForMaskedLM is a separate port of ESM-1b/ESM-2: same weights -> same logits."""
    cfg = tiny_config(arch, layers=3, embed_dim=64, heads=4, ffn_dim=128)
    sd = synthetic_state_dict(cfg, 5)
    tok = torch.randint(4, 24, (3, 19), generator=torch.Generator().manual_seed(0))
    tok[:, 0], tok[:, -1] = 0, 2
    tok[0, 3:7] = 32
    tok[1, 1:-1] = 32
    want = _hf_model(cfg, sd, arch == "esm2")(input_ids=tok, attention_mask=torch.ones_like(tok)).logits
    got = OracleModel(cfg, sd).model(tok)["logits"]
    assert (got - want).abs().max().item() < 2e-5 * want.abs().max().item()


def test_sampler_tail_identity_against_torch_categorical():
    from torch.distributions.categorical import Categorical
    for seed in range(200):
        k = [20, 3, 21, 1, 5][seed % 5]
        v = torch.randn(k, generator=torch.Generator().manual_seed(seed)).sort(descending=True).values
        torch.manual_seed(seed)
        a = int(Categorical(logits=v).sample())
        torch.manual_seed(seed)
        q = torch.empty(1, k).exponential_(1)[0]
        norm = v - v.logsumexp(-1, keepdim=True)
        assert a == int(torch.argmax(torch.softmax(norm, -1) / q))


def test_bulk_noise_draw_equals_per_call_draws():
    """draw_replay_noise relies on exponential_ consuming torch's generator element by element."""
    for k in (1, 3, 20, 21):
        torch.manual_seed(3)
        a = torch.cat([torch.empty(1, k).exponential_(1) for _ in range(257)])
        torch.manual_seed(3)
        assert torch.equal(a, torch.empty(257, k).exponential_(1))


def test_generate_step_golden(golden):
    for c in golden["generate_step"]:
        logits = torch.tensor(c["logits"])
        torch.manual_seed(c["torch_seed"])
        got = int(gibbs_loop.generate_step(logits, c["gen_idx"], temperature=c["temperature"], top_k=c["top_k"],
                                           sample=c["sample"], valid_idx=c["valid_idx"]))
        assert got == c["token"]
        valid = c["valid_idx"] if c["valid_idx"] is not None else list(range(33))
        k = effective_k(c["top_k"], len(valid), c["sample"])
        torch.manual_seed(c["torch_seed"])
        q = torch.ones(len(valid))
        q[:k] = torch.empty(1, k).exponential_(1)[0]
        assert generate_step_with_noise(logits[c["gen_idx"]], q, valid, c["top_k"], c["temperature"], c["sample"]) == c["token"]


def test_esm_generate_reproduces_reference(golden):
    for c in golden["cases"]:
        model = OracleModel(c["cfg"], synthetic_state_dict(c["cfg"], c["weights_seed"]))
        kw = {k: v for k, v in c["kwargs"].items() if k != "show_progress_bar"}
        random.seed(c["rng_seed"])
        torch.manual_seed(c["rng_seed"])
        assert gibbs_loop.esm_generate(model, **kw) == c["output"]


def test_msa_generate_reproduces_reference(golden):
    for c in golden["msa_cases"]:
        model = OracleModel(c["cfg"], synthetic_state_dict(c["cfg"], c["weights_seed"]))
        kw = {k: v for k, v in c["kwargs"].items() if k != "show_progress_bar"}
        random.seed(c["rng_seed"])
        torch.manual_seed(c["rng_seed"])
        assert gibbs_loop.msa_generate(model, **kw) == c["output"]
    for c in golden["single_cases"]:
        model = OracleModel(c["cfg"], synthetic_state_dict(c["cfg"], c["weights_seed"]))
        random.seed(c["rng_seed"])
        torch.manual_seed(c["rng_seed"])
        assert gibbs_loop.msa_generate_single(model, **c["kwargs"]) == c["output"]


def test_partition_golden(golden):
    for c in golden["fixtures"]["partition"]:
        assert gibbs_loop.partition(list(range(c["n"])), c["k"]) == c["out"]


# ------------------------------------------------------------------ second witness for the MSA / ESM-1 forwards
@pytest.mark.parametrize("shape,row_chunk,col_chunk", [((2, 4, 11), 3, 5), ((1, 1, 9), 1, 2), ((1, 7, 14), 2, 14), ((1, 5, 6), 8, 1)])
def test_msa_oracle_agrees_with_independent_witness(shape, row_chunk, col_chunk):
    """oracle/fair_esm.MSAOracle (batched float32 tensor algebra) against oracle/witness.msa_forward (numpy float64,
    explicit loops over MSA / head / row / column, tied-row scores accumulated row chunk by row chunk and columns
    walked in chunks, as fair-esm does above max_tokens_per_msa): two restatements written in different styles from the
    same published algorithm must agree to float32 rounding -- the pin for a forward no third-party port exists for."""
    import numpy as np
    from oracle.fair_esm import OracleModel
    from oracle.witness import msa_forward
    cfg = tiny_config("msa_transformer", 2, 48, 3, 96)
    sd = synthetic_state_dict(cfg, 13)
    g = torch.Generator().manual_seed(shape[-1])
    tok = torch.randint(4, 24, shape, generator=g)
    tok[..., 0] = 0
    tok[0, 0, 2:4] = 32
    tok[0, -1, 1] = 30
    got = OracleModel(cfg, sd).model(tok)["logits"].double().numpy()
    want = msa_forward(cfg, sd, tok.numpy(), row_chunk=row_chunk, col_chunk=col_chunk)
    assert got.shape == want.shape
    assert np.abs(got - want).max() / np.abs(want).max() < 2e-6


def test_esm1_oracle_agrees_with_independent_witness():
    import numpy as np
    from oracle.fair_esm import OracleModel
    from oracle.witness import esm1_forward
    cfg = tiny_config("esm1", 2, 48, 3, 96)
    sd = synthetic_state_dict(cfg, 17)
    tok = torch.randint(4, 24, (2, 13), generator=torch.Generator().manual_seed(3))
    tok[:, 0] = 32
    tok[1, 4:7] = 33
    got = OracleModel(cfg, sd).model(tok)["logits"].double().numpy()
    want = esm1_forward(cfg, sd, tok.numpy())
    assert np.abs(got - want).max() / np.abs(want).max() < 2e-6
