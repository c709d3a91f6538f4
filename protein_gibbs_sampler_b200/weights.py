"""Seeded synthetic weights in fair-esm state-dict naming.

No pretrained checkpoint exists offline, so every model here is random-init:
N(0, 0.02) linears/embeddings, LayerNorm weight 1 + small noise, small biases
(SURVEY.md section 8d).  Keys match ``esm.pretrained`` state dicts after fair-esm's
``upgrade_state_dict`` so a real checkpoint's ``model`` dict loads the same way.
"""
import torch


def synthetic_state_dict(cfg, seed=0, std=0.02, ln_noise=0.02, bias_std=0.02):
    g = torch.Generator().manual_seed(seed)
    d, F, V, L = cfg["embed_dim"], cfg["ffn_dim"], cfg["vocab"], cfg["layers"]
    sd = {}

    def lin(prefix, n_out, n_in):
        sd[prefix + ".weight"] = torch.randn(n_out, n_in, generator=g) * std
        sd[prefix + ".bias"] = torch.randn(n_out, generator=g) * bias_std

    def ln(prefix, n=d):
        sd[prefix + ".weight"] = 1.0 + torch.randn(n, generator=g) * ln_noise
        sd[prefix + ".bias"] = torch.randn(n, generator=g) * ln_noise

    sd["embed_tokens.weight"] = torch.randn(V, d, generator=g) * std * 5
    if cfg["positions"] == "learned":
        tbl = torch.randn(cfg["max_positions"] + 2, d, generator=g) * std * 5
        tbl[1].zero_()  # padding_idx row
        sd["embed_positions.weight"] = tbl
    if cfg["arch"] == "msa_transformer":
        sd["msa_position_embedding"] = torch.randn(1, 1024, 1, d, generator=g) * 0.01 * 5
    if cfg["emb_layer_norm_before"]:
        ln("emb_layer_norm_before")
    for i in range(L):
        p = "layers.%d." % i
        if cfg["arch"] == "msa_transformer":
            for blk in ("row_self_attention", "column_self_attention"):
                for nm in ("k_proj", "v_proj", "q_proj", "out_proj"):
                    lin(p + blk + ".layer." + nm, d, d)
                ln(p + blk + ".layer_norm")
            lin(p + "feed_forward_layer.layer.fc1", F, d)
            lin(p + "feed_forward_layer.layer.fc2", d, F)
            ln(p + "feed_forward_layer.layer_norm")
        else:
            for nm in ("k_proj", "v_proj", "q_proj", "out_proj"):
                lin(p + "self_attn." + nm, d, d)
            if cfg["arch"] == "esm1":  # add_bias_kv=True: one learned extra key / value per layer
                sd[p + "self_attn.bias_k"] = torch.randn(1, 1, d, generator=g) * std * 5
                sd[p + "self_attn.bias_v"] = torch.randn(1, 1, d, generator=g) * std * 5
            ln(p + "self_attn_layer_norm")
            lin(p + "fc1", F, d)
            lin(p + "fc2", d, F)
            ln(p + "final_layer_norm")
    if cfg["arch"] == "esm1":  # untied output projection straight from the last layer
        sd["embed_out"] = torch.randn(V, d, generator=g) * std * 5
        sd["embed_out_bias"] = torch.randn(V, generator=g) * bias_std
        return sd
    ln("emb_layer_norm_after")
    lin("lm_head.dense", d, d)
    ln("lm_head.layer_norm")
    sd["lm_head.weight"] = sd["embed_tokens.weight"]  # tied (RobertaLMHead)
    sd["lm_head.bias"] = torch.randn(V, generator=g) * bias_std
    return sd


def state_dict_layout(cfg):
    """[(key, shape)] of the engine-side state dict of ``cfg`` in a fixed order, derived from the geometry alone: every
    rank of a job can lay out the packed weight blob (parallel.broadcast_weights) without exchanging metadata."""
    d, F, V, L = cfg["embed_dim"], cfg["ffn_dim"], cfg["vocab"], cfg["layers"]
    out = [("embed_tokens.weight", (V, d))]

    def lin(prefix, n_out, n_in):
        out.append((prefix + ".weight", (n_out, n_in)))
        out.append((prefix + ".bias", (n_out,)))

    def ln(prefix):
        out.append((prefix + ".weight", (d,)))
        out.append((prefix + ".bias", (d,)))

    if cfg["positions"] == "learned":
        out.append(("embed_positions.weight", (cfg["max_positions"] + 2, d)))
    if cfg["arch"] == "msa_transformer":
        out.append(("msa_position_embedding", (1, 1024, 1, d)))
    if cfg["emb_layer_norm_before"]:
        ln("emb_layer_norm_before")
    for i in range(L):
        p = "layers.%d." % i
        if cfg["arch"] == "msa_transformer":
            for blk in ("row_self_attention", "column_self_attention"):
                for nm in ("k_proj", "v_proj", "q_proj", "out_proj"):
                    lin(p + blk + ".layer." + nm, d, d)
                ln(p + blk + ".layer_norm")
            lin(p + "feed_forward_layer.layer.fc1", F, d)
            lin(p + "feed_forward_layer.layer.fc2", d, F)
            ln(p + "feed_forward_layer.layer_norm")
        else:
            for nm in ("k_proj", "v_proj", "q_proj", "out_proj"):
                lin(p + "self_attn." + nm, d, d)
            if cfg["arch"] == "esm1":
                out.append((p + "self_attn.bias_k", (1, 1, d)))
                out.append((p + "self_attn.bias_v", (1, 1, d)))
            ln(p + "self_attn_layer_norm")
            lin(p + "fc1", F, d)
            lin(p + "fc2", d, F)
            ln(p + "final_layer_norm")
    if cfg["arch"] == "esm1":
        return out + [("embed_out", (V, d)), ("embed_out_bias", (V,))]
    ln("emb_layer_norm_after")
    lin("lm_head.dense", d, d)
    ln("lm_head.layer_norm")
    out.append(("lm_head.bias", (V,)))   # lm_head.weight is tied to embed_tokens.weight and never shipped
    return out


def is_gemm_weight(key):
    """The tensors the engine feeds to the tensor cores as fp16 (everything else stays fp32 on the device)."""
    return key.endswith((".q_proj.weight", ".k_proj.weight", ".v_proj.weight", ".out_proj.weight", "fc1.weight",
                         "fc2.weight", "lm_head.dense.weight"))


def sinusoidal_table(num_embeddings, dim, padding_idx=1):
    """fair-esm SinusoidalPositionalEmbedding.get_embedding (esm/modules.py), float32 like the original: row p is
    [sin(p f_0..) | cos(p f_0..)], f_j = exp(-j ln(10000) / (dim/2 - 1)); the padding row is zero."""
    import math
    half = dim // 2
    step = math.log(10000) / (half - 1)
    freq = torch.exp(torch.arange(half, dtype=torch.float) * -step)
    ang = torch.arange(num_embeddings, dtype=torch.float).unsqueeze(1) * freq.unsqueeze(0)
    tbl = torch.cat([torch.sin(ang), torch.cos(ang)], dim=1).view(num_embeddings, -1)
    if dim % 2 == 1:
        tbl = torch.cat([tbl, torch.zeros(num_embeddings, 1)], dim=1)
    tbl[padding_idx, :] = 0
    return tbl
