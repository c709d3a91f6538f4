"""Model geometry for the model families on the Gibbs hot path.

Names follow fair-esm's ``esm.pretrained.*`` loaders, which the reference binds
in ``src/pgen/models.py:59-88`` (esm1b / esm1v / esm_msa1b); the ESM-2 entries are
the additions BASELINE.json's configs 1 and 4 need.
"""

_ROBERTA = dict(arch="roberta_large", positions="learned", max_positions=1024,
                token_dropout=True, emb_layer_norm_before=True, vocab=33)
_ESM2 = dict(arch="esm2", positions="rotary", max_positions=1024,
             token_dropout=True, emb_layer_norm_before=False, vocab=33)
# ESM-1 (esm1_t6/t12/t34: the reference's esm6 / esm12 / esm34, models.py:69-82): sinusoidal positions, embeddings
# scaled by sqrt(d), one learned bias key/value per layer, LayerNorm eps 1e-12, no embedding LayerNorms, untied
# output projection, 35-token alphabet (<cls>=32, <mask>=33).
_ESM1 = dict(arch="esm1", positions="sinusoidal", max_positions=1024,
             token_dropout=False, emb_layer_norm_before=False, vocab=35)
_MSA = dict(arch="msa_transformer", positions="learned", max_positions=1024,
            token_dropout=False, emb_layer_norm_before=True, vocab=33)

MODEL_CONFIGS = {
    "esm1b_t33_650M_UR50S": dict(_ROBERTA, layers=33, embed_dim=1280, heads=20, ffn_dim=5120),
    "esm1v_t33_650M_UR90S_1": dict(_ROBERTA, layers=33, embed_dim=1280, heads=20, ffn_dim=5120),
    "esm1_t6_43M_UR50S": dict(_ESM1, layers=6, embed_dim=768, heads=12, ffn_dim=3072),
    "esm1_t12_85M_UR50S": dict(_ESM1, layers=12, embed_dim=768, heads=12, ffn_dim=3072),
    "esm1_t34_670M_UR50S": dict(_ESM1, layers=34, embed_dim=1280, heads=20, ffn_dim=5120),
    "esm2_t6_8M_UR50D": dict(_ESM2, layers=6, embed_dim=320, heads=20, ffn_dim=1280),
    "esm2_t30_150M_UR50D": dict(_ESM2, layers=30, embed_dim=640, heads=20, ffn_dim=2560),
    "esm2_t33_650M_UR50D": dict(_ESM2, layers=33, embed_dim=1280, heads=20, ffn_dim=5120),
    "esm_msa1b_t12_100M_UR50S": dict(_MSA, layers=12, embed_dim=768, heads=12, ffn_dim=3072),
}


def get_config(name, **overrides):
    cfg = dict(MODEL_CONFIGS[name])
    cfg["name"] = name
    cfg.update(overrides)
    assert cfg["embed_dim"] % cfg["heads"] == 0
    return cfg


def tiny_config(arch="esm2", layers=2, embed_dim=128, heads=2, ffn_dim=256):
    """Small geometry for parity tests the CPU oracle finishes in seconds."""
    base = {"esm2": _ESM2, "roberta_large": _ROBERTA, "msa_transformer": _MSA, "esm1": _ESM1}[arch]
    return dict(base, layers=layers, embed_dim=embed_dim, heads=heads, ffn_dim=ffn_dim,
                name="tiny_%s_l%d_d%d_h%d" % (arch, layers, embed_dim, heads))
