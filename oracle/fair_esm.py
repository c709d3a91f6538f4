"""fp32 CPU restatement of the fair-esm pieces the reference calls into.

TEST INFRASTRUCTURE (see ``oracle/__init__.py``).  ``fair-esm`` is the
third-party dependency the reference imports at ``src/pgen/models.py:1``
(``git+https://github.com/facebookresearch/esm.git``, unpinned HEAD,
``conda_env.yml:21``; last release v2.0.0).  Its source is not vendored in
``/root/reference`` and it cannot be installed offline, so its published
algorithm is restated here.  Reference call sites this file serves:

  * ``model.model(batch)["logits"]``      esm_sampler.py:223, esm_msa_sampler.py:136,236
  * ``model.alphabet.get_idx/get_tok/mask_idx/prepend_bos/append_eos``
                                          esm_sampler.py:82,92,237,262; esm_msa_sampler.py:65-66,73,259,264
  * ``model.batch_converter(batch)``      esm_sampler.py:125; esm_msa_sampler.py:89 (patched at models.py:18-56)

Everything is plain eager PyTorch in float32; no fused kernels, no autocast.
"""
import math
from typing import Dict, List, Sequence, Tuple

import torch
import torch.nn.functional as F

# --------------------------------------------------------------------------
# Alphabet / batch converters  (fair-esm esm/constants.py, esm/data.py)
# --------------------------------------------------------------------------

PROTEINSEQ_STANDARD_TOKS = list("LAGVSERTIDPKQNFYMHWCXBUZO.-")


class Alphabet:
    """Token table built the way ``esm.data.Alphabet.__init__`` builds it."""

    def __init__(self, standard_toks, prepend_toks, append_toks, prepend_bos, append_eos, use_msa):
        self.standard_toks = list(standard_toks)
        self.prepend_toks = list(prepend_toks)
        self.append_toks = list(append_toks)
        self.prepend_bos = prepend_bos
        self.append_eos = append_eos
        self.use_msa = use_msa
        self.all_toks = list(self.prepend_toks) + list(self.standard_toks)
        for i in range((8 - (len(self.all_toks) % 8)) % 8):
            self.all_toks.append("<null_%d>" % (i + 1))
        self.all_toks.extend(self.append_toks)
        self.tok_to_idx = {t: i for i, t in enumerate(self.all_toks)}
        self.unk_idx = self.tok_to_idx["<unk>"]
        self.padding_idx = self.get_idx("<pad>")
        self.cls_idx = self.get_idx("<cls>")
        self.mask_idx = self.get_idx("<mask>")
        self.eos_idx = self.get_idx("<eos>")
        self.all_special_tokens = ["<eos>", "<unk>", "<pad>", "<cls>", "<mask>"]

    def __len__(self):
        return len(self.all_toks)

    def get_idx(self, tok):
        return self.tok_to_idx.get(tok, self.unk_idx)

    def get_tok(self, ind):
        return self.all_toks[int(ind)]

    @classmethod
    def from_architecture(cls, name):
        if name in ("ESM-1", "protein_bert_base"):
            return cls(PROTEINSEQ_STANDARD_TOKS, ("<null_0>", "<pad>", "<eos>", "<unk>"),
                       ("<cls>", "<mask>", "<sep>"), True, False, False)
        if name in ("ESM-1b", "roberta_large"):
            return cls(PROTEINSEQ_STANDARD_TOKS, ("<cls>", "<pad>", "<eos>", "<unk>"),
                       ("<mask>",), True, True, False)
        if name in ("MSA Transformer", "msa_transformer"):
            return cls(PROTEINSEQ_STANDARD_TOKS, ("<cls>", "<pad>", "<eos>", "<unk>"),
                       ("<mask>",), True, False, True)
        raise ValueError("Unknown architecture selected")

    def tokenize(self, text):
        """Split on literal special tokens (``<mask>`` is ONE token), else per char."""
        out = []
        i = 0
        specials = sorted((t for t in self.all_toks if len(t) > 1), key=len, reverse=True)
        while i < len(text):
            if text[i] == "<":
                hit = next((t for t in specials if text.startswith(t, i)), None)
                if hit is not None:
                    out.append(hit)
                    i += len(hit)
                    continue
            if not text[i].isspace():
                out.append(text[i])
            i += 1
        return out

    def encode(self, text):
        return [self.get_idx(t) for t in self.tokenize(text)]

    def get_batch_converter(self):
        return MSABatchConverter(self) if self.use_msa else BatchConverter(self)


class BatchConverter:
    """``esm.data.BatchConverter.__call__``: (label, str) list -> int64 (B, maxlen+bos+eos)."""

    def __init__(self, alphabet):
        self.alphabet = alphabet

    def __call__(self, raw_batch):
        a = self.alphabet
        labels = [l for l, _ in raw_batch]
        strs = [s for _, s in raw_batch]
        enc = [a.encode(s) for s in strs]
        max_len = max(len(e) for e in enc)
        tokens = torch.full((len(raw_batch), max_len + int(a.prepend_bos) + int(a.append_eos)),
                            a.padding_idx, dtype=torch.int64)
        for i, e in enumerate(enc):
            if a.prepend_bos:
                tokens[i, 0] = a.cls_idx
            tokens[i, int(a.prepend_bos):len(e) + int(a.prepend_bos)] = torch.tensor(e, dtype=torch.int64)
            if a.append_eos:
                tokens[i, len(e) + int(a.prepend_bos)] = a.eos_idx
        return labels, strs, tokens


def rawbatchlen(raw):
    """Length where a ``<...>`` literal counts as one column (reference models.py:6-16)."""
    n, counting = 0, True
    for ch in raw:
        if ch == "<":
            counting = False
        if ch == ">":
            counting = True
        if counting:
            n += 1
    return n


class MSABatchConverter(BatchConverter):
    """The reference's patched MSA converter (models.py:18-54)."""

    def __call__(self, inputs):
        raw_batch = [inputs] if isinstance(inputs[0][0], str) else inputs
        a = self.alphabet
        rows = max(len(m) for m in raw_batch)
        cols = max(rawbatchlen(m[0][1]) for m in raw_batch)
        tokens = torch.full((len(raw_batch), rows, cols + int(a.prepend_bos) + int(a.append_eos)),
                            a.padding_idx, dtype=torch.int64)
        labels, strs = [], []
        for i, msa in enumerate(raw_batch):
            if len({rawbatchlen(s) for _, s in msa}) != 1:
                raise RuntimeError("Received unaligned sequences for input to MSA, all sequence "
                                   "lengths must be equal.")
            l, s, t = BatchConverter.__call__(self, msa)
            labels.append(l)
            strs.append(s)
            tokens[i, :t.size(0), :t.size(1)] = t
        return labels, strs, tokens


# --------------------------------------------------------------------------
# Shared numerics
# --------------------------------------------------------------------------

def gelu_erf(x):
    return x * 0.5 * (1.0 + torch.erf(x / math.sqrt(2.0)))


def layer_norm(x, w, b, eps):
    return F.layer_norm(x, (x.shape[-1],), w, b, eps)


def linear(x, sd, prefix):
    return F.linear(x, sd[prefix + ".weight"], sd[prefix + ".bias"])


def rotate_half(x):
    x1, x2 = x.chunk(2, dim=-1)
    return torch.cat((-x2, x1), dim=-1)


def rotary(x, dim):
    """fair-esm esm/rotary_embedding.py: x is (B*H, T, Dh)."""
    t_len = x.shape[-2]
    inv_freq = 1.0 / (10000 ** (torch.arange(0, dim, 2).float() / dim))
    t = torch.arange(t_len).type_as(inv_freq)
    freqs = torch.einsum("i,j->ij", t, inv_freq)
    emb = torch.cat((freqs, freqs), dim=-1)
    cos, sin = emb.cos()[None], emb.sin()[None]
    return x * cos + rotate_half(x) * sin


def roberta_lm_head(x, sd, eps=1e-5):
    x = linear(x, sd, "lm_head.dense")
    x = gelu_erf(x)
    x = layer_norm(x, sd["lm_head.layer_norm.weight"], sd["lm_head.layer_norm.bias"], eps)
    return F.linear(x, sd["lm_head.weight"]) + sd["lm_head.bias"]


def learned_positions(tokens, table, padding_idx):
    mask = tokens.ne(padding_idx).int()
    pos = (torch.cumsum(mask, dim=1).type_as(mask) * mask).long() + padding_idx
    return F.embedding(pos, table, padding_idx)


# --------------------------------------------------------------------------
# Single-sequence models: ESM-1b / ESM-1v (learned positions) and ESM-2 (rotary)
# --------------------------------------------------------------------------

class ESMOracle(torch.nn.Module):
    """ProteinBertModel (roberta_large arch) / ESM2 forward, eval mode."""

    def __init__(self, cfg: dict, state_dict: Dict[str, torch.Tensor], hook=None):
        super().__init__()
        self.cfg = dict(cfg)
        self.sd = {k: v.detach().float().clone() for k, v in state_dict.items()}
        self.hook = hook  # optional callable(name, tensor) for per-stage parity taps
        self.alphabet = Alphabet.from_architecture("ESM-1b")
        self.mm = torch.matmul  # replaced by precision-emulation experiments in tests

    def to(self, *a, **k):  # the sampler calls model.to(device); CPU only
        return self

    def _tap(self, name, t):
        if self.hook is not None:
            self.hook(name, t)

    def _lin(self, x, prefix):
        return self.mm(x, self.sd[prefix + ".weight"].t()) + self.sd[prefix + ".bias"]

    def forward(self, tokens, repr_layers=(), **kw):
        cfg, sd, a = self.cfg, self.sd, self.alphabet
        d, H, eps = cfg["embed_dim"], cfg["heads"], 1e-5
        Dh = d // H
        B, T = tokens.shape
        pad = tokens.eq(a.padding_idx)
        x = F.embedding(tokens, sd["embed_tokens.weight"])
        if cfg.get("token_dropout", True):
            x = x.masked_fill((tokens == a.mask_idx).unsqueeze(-1), 0.0)
            mask_ratio_train = 0.15 * 0.8
            src_lengths = (~pad).sum(-1)
            mask_ratio_observed = (tokens == a.mask_idx).sum(-1).to(x.dtype) / src_lengths
            x = x * (1 - mask_ratio_train) / (1 - mask_ratio_observed)[:, None, None]
        if cfg["positions"] == "learned":
            x = x + learned_positions(tokens, sd["embed_positions.weight"], a.padding_idx)
            x = layer_norm(x, sd["emb_layer_norm_before.weight"], sd["emb_layer_norm_before.bias"], eps)
        x = x * (1 - pad.unsqueeze(-1).type_as(x))
        self._tap("embed", x)
        key_bias = None
        if pad.any():
            key_bias = torch.zeros(B, 1, 1, T).masked_fill(pad[:, None, None, :], float("-inf"))
        for i in range(cfg["layers"]):
            p = "layers.%d." % i
            h = layer_norm(x, sd[p + "self_attn_layer_norm.weight"], sd[p + "self_attn_layer_norm.bias"], eps)
            q = self._lin(h, p + "self_attn.q_proj") * (Dh ** -0.5)
            k = self._lin(h, p + "self_attn.k_proj")
            v = self._lin(h, p + "self_attn.v_proj")
            q = q.view(B, T, H, Dh).transpose(1, 2)
            k = k.view(B, T, H, Dh).transpose(1, 2)
            v = v.view(B, T, H, Dh).transpose(1, 2)
            if cfg["positions"] == "rotary":
                q = rotary(q.reshape(B * H, T, Dh), Dh).view(B, H, T, Dh)
                k = rotary(k.reshape(B * H, T, Dh), Dh).view(B, H, T, Dh)
            s = self.mm(q, k.transpose(-1, -2))
            if key_bias is not None:
                s = s + key_bias
            pr = torch.softmax(s, dim=-1, dtype=torch.float64 if s.dtype == torch.float64 else torch.float32)
            ctx = self.mm(pr, v).transpose(1, 2).reshape(B, T, d)
            x = x + self._lin(ctx, p + "self_attn.out_proj")
            self._tap("attn%d" % i, x)
            h = layer_norm(x, sd[p + "final_layer_norm.weight"], sd[p + "final_layer_norm.bias"], eps)
            h = gelu_erf(self._lin(h, p + "fc1"))
            x = x + self._lin(h, p + "fc2")
            self._tap("layer%d" % i, x)
        x = layer_norm(x, sd["emb_layer_norm_after.weight"], sd["emb_layer_norm_after.bias"], eps)
        self._tap("final_ln", x)
        h = gelu_erf(self._lin(x, "lm_head.dense"))
        h = layer_norm(h, sd["lm_head.layer_norm.weight"], sd["lm_head.layer_norm.bias"], eps)
        logits = torch.matmul(h, sd["lm_head.weight"].t()) + sd["lm_head.bias"]
        return {"logits": logits, "representations": {}}


# --------------------------------------------------------------------------
# ESM-1 (esm1_t6 / t12 / t34; ProteinBertModel with arch != roberta_large, esm/model/esm1.py)
# --------------------------------------------------------------------------

def sinusoidal_positions(tokens, dim, padding_idx):
    """esm/modules.py SinusoidalPositionalEmbedding: table row = position + padding_idx + 1 for non-pad tokens."""
    B, T = tokens.shape
    n = padding_idx + 1 + T
    half = dim // 2
    step = math.log(10000) / (half - 1)
    freq = torch.exp(torch.arange(half, dtype=torch.float) * -step)
    ang = torch.arange(n, dtype=torch.float).unsqueeze(1) * freq.unsqueeze(0)
    tbl = torch.cat([torch.sin(ang), torch.cos(ang)], dim=1).view(n, -1)
    if dim % 2 == 1:
        tbl = torch.cat([tbl, torch.zeros(n, 1)], dim=1)
    tbl[padding_idx, :] = 0
    mask = tokens.ne(padding_idx)
    pos = (torch.arange(T) + padding_idx + 1).expand_as(tokens) * mask.long() + padding_idx * (1 - mask.long())
    return tbl.index_select(0, pos.reshape(-1)).view(B, T, -1)


class ESM1Oracle(torch.nn.Module):
    """embed_scale sqrt(d), sinusoidal positions, no embedding LayerNorms / token dropout, pre-LN layers whose
    attention has one learned bias key/value appended after the sequence (add_bias_kv), LayerNorm eps 1e-12
    (ESM1LayerNorm), logits = F.linear(x, embed_out, embed_out_bias)."""

    def __init__(self, cfg: dict, state_dict: Dict[str, torch.Tensor], hook=None):
        super().__init__()
        self.cfg = dict(cfg)
        self.sd = {k: v.detach().float().clone() for k, v in state_dict.items()}
        self.hook = hook
        self.alphabet = Alphabet.from_architecture("ESM-1")
        self.mm = torch.matmul

    def to(self, *a, **k):
        return self

    def _lin(self, x, prefix):
        return self.mm(x, self.sd[prefix + ".weight"].t()) + self.sd[prefix + ".bias"]

    def forward(self, tokens, repr_layers=(), **kw):
        cfg, sd, a = self.cfg, self.sd, self.alphabet
        d, H, eps = cfg["embed_dim"], cfg["heads"], 1e-12
        Dh = d // H
        B, T = tokens.shape
        pad = tokens.eq(a.padding_idx)
        x = math.sqrt(d) * F.embedding(tokens, sd["embed_tokens.weight"])
        x = x + sinusoidal_positions(tokens, d, a.padding_idx)
        x = x * (1 - pad.unsqueeze(-1).type_as(x))
        if self.hook is not None:
            self.hook("embed", x)
        for i in range(cfg["layers"]):
            p = "layers.%d." % i
            h = layer_norm(x, sd[p + "self_attn_layer_norm.weight"], sd[p + "self_attn_layer_norm.bias"], eps)
            q = self._lin(h, p + "self_attn.q_proj") * (Dh ** -0.5)
            k = torch.cat([self._lin(h, p + "self_attn.k_proj"), sd[p + "self_attn.bias_k"].view(1, 1, d).expand(B, 1, d)], 1)
            v = torch.cat([self._lin(h, p + "self_attn.v_proj"), sd[p + "self_attn.bias_v"].view(1, 1, d).expand(B, 1, d)], 1)
            q = q.view(B, T, H, Dh).transpose(1, 2)
            k = k.view(B, T + 1, H, Dh).transpose(1, 2)
            v = v.view(B, T + 1, H, Dh).transpose(1, 2)
            s = self.mm(q, k.transpose(-1, -2))
            if pad.any():
                s = s.masked_fill(torch.cat([pad, pad.new_zeros(B, 1)], 1)[:, None, None, :], float("-inf"))
            pr = torch.softmax(s, dim=-1, dtype=torch.float32)
            ctx = self.mm(pr, v).transpose(1, 2).reshape(B, T, d)
            x = x + self._lin(ctx, p + "self_attn.out_proj")
            h = layer_norm(x, sd[p + "final_layer_norm.weight"], sd[p + "final_layer_norm.bias"], eps)
            x = x + self._lin(gelu_erf(self._lin(h, p + "fc1")), p + "fc2")
            if self.hook is not None:
                self.hook("layer%d" % i, x)
        bias = sd.get("embed_out_bias")
        return {"logits": F.linear(x, sd["embed_out"], bias), "representations": {}}


# --------------------------------------------------------------------------
# MSA Transformer (esm/model/msa_transformer.py, esm/axial_attention.py)
# --------------------------------------------------------------------------

class MSAOracle(torch.nn.Module):
    def __init__(self, cfg: dict, state_dict: Dict[str, torch.Tensor], hook=None):
        super().__init__()
        self.cfg = dict(cfg)
        self.sd = {k: v.detach().float().clone() for k, v in state_dict.items()}
        self.hook = hook
        self.alphabet = Alphabet.from_architecture("MSA Transformer")

    def to(self, *a, **k):
        return self

    def _tap(self, name, t):
        if self.hook is not None:
            self.hook(name, t)

    def forward(self, tokens, **kw):
        cfg, sd, a = self.cfg, self.sd, self.alphabet
        d, H, eps = cfg["embed_dim"], cfg["heads"], 1e-5
        Dh = d // H
        assert tokens.ndim == 3
        B, R, C = tokens.shape
        pad = tokens.eq(a.padding_idx)
        x = F.embedding(tokens, sd["embed_tokens.weight"])
        x = x + learned_positions(tokens.view(B * R, C), sd["embed_positions.weight"], a.padding_idx).view(x.size())
        if R > 1024:
            raise RuntimeError("Using model with MSA position embedding trained on maximum MSA depth of 1024, "
                               "but received %d alignments." % R)
        x = x + sd["msa_position_embedding"][:, :R]
        x = layer_norm(x, sd["emb_layer_norm_before.weight"], sd["emb_layer_norm_before.bias"], eps)
        x = x * (1 - pad.unsqueeze(-1).type_as(x))
        self._tap("embed", x)
        x = x.permute(1, 2, 0, 3)  # R, C, B, D
        for i in range(cfg["layers"]):
            p = "layers.%d." % i
            # --- tied row attention
            pp = p + "row_self_attention."
            h = layer_norm(x, sd[pp + "layer_norm.weight"], sd[pp + "layer_norm.bias"], eps)
            q = linear(h, sd, pp + "layer.q_proj").view(R, C, B, H, Dh) * ((Dh ** -0.5) / math.sqrt(R))
            k = linear(h, sd, pp + "layer.k_proj").view(R, C, B, H, Dh)
            v = linear(h, sd, pp + "layer.v_proj").view(R, C, B, H, Dh)
            if pad.any():
                q = q * (1 - pad.permute(1, 2, 0).unsqueeze(3).unsqueeze(4).to(q))
            s = torch.einsum("rinhd,rjnhd->hnij", q, k)
            if pad.any():
                s = s.masked_fill(pad[:, 0].unsqueeze(0).unsqueeze(2), -10000)
            pr = torch.softmax(s, dim=-1)
            ctx = torch.einsum("hnij,rjnhd->rinhd", pr, v).reshape(R, C, B, d)
            x = x + linear(ctx, sd, pp + "layer.out_proj")
            self._tap("row%d" % i, x)
            # --- column attention
            pp = p + "column_self_attention."
            h = layer_norm(x, sd[pp + "layer_norm.weight"], sd[pp + "layer_norm.bias"], eps)
            if R == 1:
                out = linear(linear(h, sd, pp + "layer.v_proj"), sd, pp + "layer.out_proj")
            else:
                q = linear(h, sd, pp + "layer.q_proj").view(R, C, B, H, Dh) * (Dh ** -0.5)
                k = linear(h, sd, pp + "layer.k_proj").view(R, C, B, H, Dh)
                v = linear(h, sd, pp + "layer.v_proj").view(R, C, B, H, Dh)
                s = torch.einsum("icnhd,jcnhd->hcnij", q, k)
                if pad.any():
                    s = s.masked_fill(pad.permute(2, 0, 1).unsqueeze(0).unsqueeze(3), -10000)
                pr = torch.softmax(s, dim=-1)
                ctx = torch.einsum("hcnij,jcnhd->icnhd", pr, v).reshape(R, C, B, d)
                out = linear(ctx, sd, pp + "layer.out_proj")
            x = x + out
            self._tap("col%d" % i, x)
            # --- feed forward
            pp = p + "feed_forward_layer."
            h = layer_norm(x, sd[pp + "layer_norm.weight"], sd[pp + "layer_norm.bias"], eps)
            h = F.gelu(linear(h, sd, pp + "layer.fc1"))
            x = x + linear(h, sd, pp + "layer.fc2")
            self._tap("layer%d" % i, x)
        x = layer_norm(x, sd["emb_layer_norm_after.weight"], sd["emb_layer_norm_after.bias"], eps)
        x = x.permute(2, 0, 1, 3)  # B, R, C, D
        return {"logits": roberta_lm_head(x, sd, eps), "representations": {}}


class OracleModel:
    """The duck-typed triple the reference samplers expect (esm_sampler.py:55-57)."""

    def __init__(self, cfg, state_dict, hook=None):
        if cfg["arch"] == "msa_transformer":
            self.model = MSAOracle(cfg, state_dict, hook)
        elif cfg["arch"] == "esm1":
            self.model = ESM1Oracle(cfg, state_dict, hook)
        else:
            self.model = ESMOracle(cfg, state_dict, hook)
        self.alphabet = self.model.alphabet
        self.batch_converter = self.alphabet.get_batch_converter()
