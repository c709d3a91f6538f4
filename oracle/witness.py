"""Second, structurally different CPU witness for the two forwards no third-party port exists for: the MSA Transformer
and the ESM-1 family (TEST INFRASTRUCTURE, see oracle/__init__.py; only tests/ may import it).

`oracle/fair_esm.py` restates fair-esm with batched torch tensor algebra (einsum over all rows / heads at once, torch's
fused layer_norm / softmax / gelu, float32).  This module follows the same PUBLISHED algorithm (SURVEY.md Appendix A.2 /
A.4: fair-esm esm/model/msa_transformer.py, esm/axial_attention.py, esm/multihead_attention.py, esm/modules.py) in a
deliberately different style, so that an indexing or transposition slip in one of them cannot hide in the other:

  * numpy, float64 throughout (differences against the float32 oracle are rounding, ~1e-6);
  * explicit Python loops over MSA / head / alignment row / column instead of einsum; every primitive (LayerNorm,
    softmax, erf-GELU, position tables) written out from its definition;
  * tied row attention accumulates the scores ROW CHUNK BY ROW CHUNK before the softmax and column attention walks the
    columns in chunks -- the schedule fair-esm itself switches to above `max_tokens_per_msa` -- so the chunked and the
    one-shot forms are checked against each other as well.

The reference reaches these forwards at /root/reference/src/pgen/esm_msa_sampler.py:136,236 (MSA) and
esm_sampler.py:223 with models.ESM6 / ESM12 / ESM34 (ESM-1).
"""
import math

import numpy as np


def _np(t):
    return np.asarray(t.detach().cpu().numpy() if hasattr(t, "detach") else t, dtype=np.float64)


_erf = np.vectorize(math.erf, otypes=[np.float64])


def _gelu(x):
    return 0.5 * x * (1.0 + _erf(x / math.sqrt(2.0)))


def _layer_norm(x, w, b, eps):
    mu = x.mean(axis=-1, keepdims=True)
    var = ((x - mu) ** 2).mean(axis=-1, keepdims=True)          # biased variance, as torch
    return (x - mu) / np.sqrt(var + eps) * w + b


def _softmax_last(s):
    s = s - s.max(axis=-1, keepdims=True)
    e = np.exp(s)
    return e / e.sum(axis=-1, keepdims=True)


def _lin(x, sd, prefix):
    return x @ _np(sd[prefix + ".weight"]).T + _np(sd[prefix + ".bias"])


def msa_forward(cfg, sd, tokens, row_chunk=3, col_chunk=5):
    """MSA Transformer logits [B, R, C, V] (no <pad> on the Gibbs path: SURVEY App. B.7)."""
    tokens = np.asarray(tokens)
    B, R, C = tokens.shape
    d, H, eps = cfg["embed_dim"], cfg["heads"], 1e-5
    Dh = d // H
    E = _np(sd["embed_tokens.weight"])
    P = _np(sd["embed_positions.weight"])                       # learned, row 1 = padding, first token -> row 2
    Prow = _np(sd["msa_position_embedding"]).reshape(-1, d)     # [1024, d]
    out = np.zeros((B, R, C, cfg["vocab"]))
    for b in range(B):
        x = np.zeros((R, C, d))
        for r in range(R):
            for c in range(C):
                x[r, c] = E[tokens[b, r, c]] + P[c + 2] + Prow[r]
        x = _layer_norm(x, _np(sd["emb_layer_norm_before.weight"]), _np(sd["emb_layer_norm_before.bias"]), eps)
        for layer in range(cfg["layers"]):
            p = "layers.%d." % layer
            # ---- tied row attention: ONE attention map per head, shared by all rows; scores summed over rows
            pp = p + "row_self_attention."
            h = _layer_norm(x, _np(sd[pp + "layer_norm.weight"]), _np(sd[pp + "layer_norm.bias"]), eps)
            q = _lin(h, sd, pp + "layer.q_proj") * (Dh ** -0.5) / math.sqrt(R)
            k = _lin(h, sd, pp + "layer.k_proj")
            v = _lin(h, sd, pp + "layer.v_proj")
            ctx = np.zeros((R, C, d))
            for hd in range(H):
                sl = slice(hd * Dh, (hd + 1) * Dh)
                scores = np.zeros((C, C))
                for r0 in range(0, R, row_chunk):               # fair-esm's chunked accumulation (axial_attention.py)
                    for r in range(r0, min(r0 + row_chunk, R)):
                        scores += q[r, :, sl] @ k[r, :, sl].T
                probs = _softmax_last(scores)
                for r in range(R):
                    ctx[r, :, sl] = probs @ v[r, :, sl]
            x = x + _lin(ctx, sd, pp + "layer.out_proj")
            # ---- column attention: every alignment column attends over the R rows
            pp = p + "column_self_attention."
            h = _layer_norm(x, _np(sd[pp + "layer_norm.weight"]), _np(sd[pp + "layer_norm.bias"]), eps)
            if R == 1:                                          # fair-esm's short cut: softmax over one key is 1
                upd = _lin(_lin(h, sd, pp + "layer.v_proj"), sd, pp + "layer.out_proj")
            else:
                q = _lin(h, sd, pp + "layer.q_proj") * (Dh ** -0.5)
                k = _lin(h, sd, pp + "layer.k_proj")
                v = _lin(h, sd, pp + "layer.v_proj")
                ctx = np.zeros((R, C, d))
                for c0 in range(0, C, col_chunk):
                    for c in range(c0, min(c0 + col_chunk, C)):
                        for hd in range(H):
                            sl = slice(hd * Dh, (hd + 1) * Dh)
                            probs = _softmax_last(q[:, c, sl] @ k[:, c, sl].T)      # [R, R]
                            ctx[:, c, sl] = probs @ v[:, c, sl]
                upd = _lin(ctx, sd, pp + "layer.out_proj")
            x = x + upd
            # ---- feed forward
            pp = p + "feed_forward_layer."
            h = _layer_norm(x, _np(sd[pp + "layer_norm.weight"]), _np(sd[pp + "layer_norm.bias"]), eps)
            x = x + _lin(_gelu(_lin(h, sd, pp + "layer.fc1")), sd, pp + "layer.fc2")
        x = _layer_norm(x, _np(sd["emb_layer_norm_after.weight"]), _np(sd["emb_layer_norm_after.bias"]), eps)
        # RobertaLMHead: dense -> gelu -> LayerNorm -> tied projection + bias
        y = _gelu(_lin(x, sd, "lm_head.dense"))
        y = _layer_norm(y, _np(sd["lm_head.layer_norm.weight"]), _np(sd["lm_head.layer_norm.bias"]), eps)
        out[b] = y @ E.T + _np(sd["lm_head.bias"])
    return out


def esm1_forward(cfg, sd, tokens):
    """ESM-1 (esm1_t6 / t12 / t34) logits [B, T, V]: sqrt(d)-scaled embeddings, sinusoidal positions, one learned bias
    key / value per layer appended after the sequence, LayerNorm eps 1e-12, untied output projection."""
    tokens = np.asarray(tokens)
    B, T = tokens.shape
    d, H, eps = cfg["embed_dim"], cfg["heads"], 1e-12
    Dh = d // H
    E = _np(sd["embed_tokens.weight"])
    half = d // 2
    out = np.zeros((B, T, cfg["vocab"]))
    for b in range(B):
        x = np.zeros((T, d))
        for t in range(T):
            pos = t + 2                                         # padding_idx + 1 + t
            # fair-esm builds the table in float32 (exp / sin / cos of float32 arguments): reproduce its rounding
            freq = np.exp(np.arange(half, dtype=np.float32) * np.float32(-math.log(10000) / (half - 1)))
            ang = np.float32(pos) * freq
            x[t] = math.sqrt(d) * E[tokens[b, t]] + np.concatenate([np.sin(ang), np.cos(ang)]).astype(np.float64)
        for layer in range(cfg["layers"]):
            p = "layers.%d." % layer
            h = _layer_norm(x, _np(sd[p + "self_attn_layer_norm.weight"]), _np(sd[p + "self_attn_layer_norm.bias"]), eps)
            q = _lin(h, sd, p + "self_attn.q_proj") * (Dh ** -0.5)
            k = np.vstack([_lin(h, sd, p + "self_attn.k_proj"), _np(sd[p + "self_attn.bias_k"]).reshape(1, d)])
            v = np.vstack([_lin(h, sd, p + "self_attn.v_proj"), _np(sd[p + "self_attn.bias_v"]).reshape(1, d)])
            ctx = np.zeros((T, d))
            for hd in range(H):
                sl = slice(hd * Dh, (hd + 1) * Dh)
                for t in range(T):                              # one query at a time
                    w = _softmax_last(k[:, sl] @ q[t, sl])
                    ctx[t, sl] = w @ v[:, sl]
            x = x + _lin(ctx, sd, p + "self_attn.out_proj")
            h = _layer_norm(x, _np(sd[p + "final_layer_norm.weight"]), _np(sd[p + "final_layer_norm.bias"]), eps)
            x = x + _lin(_gelu(_lin(h, sd, p + "fc1")), sd, p + "fc2")
        out[b] = x @ _np(sd["embed_out"]).T + _np(sd["embed_out_bias"])
    return out
