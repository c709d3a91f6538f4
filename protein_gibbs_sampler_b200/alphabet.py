"""Tokenisation for the ESM model families (host side of the drop-in boundary).

Mirrors what the reference obtains from fair-esm: ``model.alphabet`` (``get_idx``, ``get_tok``, ``mask_idx``,
``prepend_bos``, ``append_eos``, ``padding_idx``; used at /root/reference/src/pgen/esm_sampler.py:82,92,237,262)
and ``model.batch_converter`` (esm_sampler.py:125), including the reference's own patched MSA converter
(/root/reference/src/pgen/models.py:6-54) whose ``<mask>`` literals count as a single column.
Token ids are pinned by the reference tests (test_esm_sampler.py:43-88, test_esm_msa_sampler.py:43-84).
"""
import re

import torch

_RESIDUES = "LAGVSERTIDPKQNFYMHWCXBUZO.-"
_SPECIAL = re.compile(r"<[a-z_0-9]+>")


class Alphabet:
    def __init__(self, prepend, append, prepend_bos, append_eos, use_msa=False):
        toks = list(prepend) + list(_RESIDUES)
        n = 1
        while len(toks) % 8:
            toks.append("<null_%d>" % n)
            n += 1
        toks += list(append)
        self.all_toks = toks
        self.standard_toks = list(_RESIDUES)
        self.prepend_toks, self.append_toks = list(prepend), list(append)
        self.tok_to_idx = {t: i for i, t in enumerate(toks)}
        self.prepend_bos, self.append_eos, self.use_msa = prepend_bos, append_eos, use_msa
        self.unk_idx = self.tok_to_idx["<unk>"]
        self.padding_idx = self.tok_to_idx["<pad>"]
        self.cls_idx = self.tok_to_idx["<cls>"]
        self.mask_idx = self.tok_to_idx["<mask>"]
        self.eos_idx = self.tok_to_idx["<eos>"]

    @classmethod
    def esm1b(cls):
        """ESM-1b / ESM-1v / ESM-2 alphabet: 33 tokens, <cls>=0 ... <mask>=32, bos and eos."""
        return cls(("<cls>", "<pad>", "<eos>", "<unk>"), ("<mask>",), True, True)

    @classmethod
    def esm1(cls):
        """ESM-1 (esm1_t6/t12/t34) alphabet: 35 tokens, <null_0>=0 ... <cls>=32, <mask>=33, <sep>=34; bos only
        (pinned by the reference's fixtures, test_esm_sampler.py:43-66)."""
        return cls(("<null_0>", "<pad>", "<eos>", "<unk>"), ("<cls>", "<mask>", "<sep>"), True, False)

    @classmethod
    def msa(cls):
        """MSA Transformer alphabet: same ids, no <eos> appended."""
        return cls(("<cls>", "<pad>", "<eos>", "<unk>"), ("<mask>",), True, False, use_msa=True)

    def __len__(self):
        return len(self.all_toks)

    def get_idx(self, tok):
        return self.tok_to_idx.get(tok, self.unk_idx)

    def get_tok(self, ind):
        return self.all_toks[int(ind)]

    def to_dict(self):
        return dict(self.tok_to_idx)

    def tokenize(self, text):
        out, pos = [], 0
        for m in _SPECIAL.finditer(text):
            if m.group(0) in self.tok_to_idx:
                out.extend(ch for ch in text[pos:m.start()] if not ch.isspace())
                out.append(m.group(0))
                pos = m.end()
        out.extend(ch for ch in text[pos:] if not ch.isspace())
        return out

    def encode(self, text):
        return [self.get_idx(t) for t in self.tokenize(text)]

    def get_batch_converter(self):
        return MSABatchConverter(self) if self.use_msa else BatchConverter(self)


class BatchConverter:
    """[(label, sequence)] -> (labels, strs, int64 tokens [B, max_len + bos + eos]) padded with <pad>."""

    def __init__(self, alphabet):
        self.alphabet = alphabet

    def __call__(self, raw_batch):
        a = self.alphabet
        labels, strs = [l for l, _ in raw_batch], [s for _, s in raw_batch]
        ids = [a.encode(s) for s in strs]
        width = max(len(x) for x in ids) + int(a.prepend_bos) + int(a.append_eos)
        tokens = torch.full((len(ids), width), a.padding_idx, dtype=torch.int64)
        for i, x in enumerate(ids):
            row = ([a.cls_idx] if a.prepend_bos else []) + x + ([a.eos_idx] if a.append_eos else [])
            tokens[i, :len(row)] = torch.tensor(row, dtype=torch.int64)
        return labels, strs, tokens


def rawbatchlen(raw):
    """Number of alignment columns in a raw string whose ``<...>`` literals are one column each."""
    return len(_SPECIAL.sub("#", raw)) if "<" in raw else len(raw)


class MSABatchConverter(BatchConverter):
    """One MSA ([(label, row)]) or a list of MSAs -> int64 tokens [B, R, C + bos]."""

    def __call__(self, inputs):
        raw_batch = [inputs] if isinstance(inputs[0][0], str) else inputs
        a = self.alphabet
        depth = max(len(m) for m in raw_batch)
        width = max(rawbatchlen(m[0][1]) for m in raw_batch) + int(a.prepend_bos) + int(a.append_eos)
        tokens = torch.full((len(raw_batch), depth, width), a.padding_idx, dtype=torch.int64)
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
