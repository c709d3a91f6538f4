"""Model triples (``.model`` / ``.alphabet`` / ``.batch_converter``) the samplers consume.

Same contract as /root/reference/src/pgen/models.py:59-88 (stated at esm_sampler.py:55-57), but ``.model``
is a handle on the CUDA engine instead of a fair-esm ``nn.Module``.  ESM1b / ESM1v / ESM6 / ESM12 / ESM34 / ESM_MSA1
keep the reference's class names; ESM2_t6_8M / ESM2_t30_150M / ESM2_t33_650M are additions (BASELINE configs 1 and 4).

No pretrained checkpoint can be downloaded in this environment: by default the weights are seeded synthetic
tensors (``weights.synthetic_state_dict``).  Pass ``checkpoint=<path to a fair-esm .pt>`` or
``state_dict=<dict>`` to run real weights.
"""
import re

import torch

from .alphabet import Alphabet
from .config import get_config
from .engine import Engine
from .weights import synthetic_state_dict


def upgrade_state_dict(sd, arch=None):
    """Key rewriting of fair-esm's checkpoint loaders (``esm/pretrained.py``: ``_load_model_and_alphabet_core_v1`` /
    ``_core_v2``), per architecture -- the reference reaches it through ``esm.pretrained.*`` at models.py:61-86.

    * ``roberta_large`` (ESM-1b / ESM-1v): drop ``encoder.`` / ``sentence_encoder.``; the ``<mask>`` row of
      ``embed_tokens.weight`` is zeroed ("for token drop") -- the tied LM head then gives ``<mask>`` the bare output
      bias, which enters every full-vocabulary log_softmax of the likelihood path.
    * ``protein_bert_base`` (ESM-1: esm1_t6 / t12 / t34): drop ``decoder.``.
    * ``msa_transformer``: the published checkpoints name the two axial attentions the other way round, so ``row`` and
      ``column`` are SWAPPED in every key before the prefixes are dropped.  Both blocks are d x d: without the swap
      the weights load without an error and every result is silently wrong.
    * ESM-2 (``_core_v2``): drop ``encoder.sentence_encoder.`` / ``encoder.`` at the start of the key.
    ``arch`` is this package's architecture name (``cfg["arch"]``); ``None`` only strips prefixes (already-upgraded
    or synthetic dicts pass through unchanged)."""
    def strip_all(key, marker):
        # fair-esm: "".join(s.split(marker)[1:] if marker[:-1] in s else s) -- everything up to and including the
        # first occurrence goes, later occurrences are removed as well
        return "".join(key.split(marker)[1:]) if marker in key else key

    out = {}
    for k, v in sd.items():
        if arch == "msa_transformer":
            k = k.replace("row", "column") if "row" in k else k.replace("column", "row")
            k = strip_all(strip_all(k, "sentence_encoder."), "encoder.")
        elif arch == "roberta_large":
            k = strip_all(strip_all(k, "sentence_encoder."), "encoder.")
        elif arch == "esm1":
            k = strip_all(k, "decoder.")
        else:  # esm2, or unknown: prefixes at the start of the key only
            k = re.sub(r"^(model\.)?(encoder\.sentence_encoder\.|encoder\.)", "", k)
        k = re.sub(r"^msa\.", "", k)
        out[k] = v
    if arch == "roberta_large" and "embed_tokens.weight" in out:
        w = out["embed_tokens.weight"].clone()
        w[32].zero_()   # alphabet.mask_idx of the "ESM-1b" alphabet
        out["embed_tokens.weight"] = w
        if "lm_head.weight" in out:
            out["lm_head.weight"] = w
    return out


def load_checkpoint(path, arch):
    """``model`` state dict of a fair-esm ``.pt`` checkpoint, upgraded for ``arch``.  Loaded with
    ``weights_only=True`` (tensors, containers and the ``argparse.Namespace`` fair-esm stores its hyper-parameters
    in): a checkpoint is user-supplied data and the full pickle protocol can run arbitrary code.  Set
    ``PGIBBS_TRUST_CHECKPOINT=1`` to fall back to the unrestricted loader for files you trust."""
    import argparse
    import os
    try:
        with torch.serialization.safe_globals([argparse.Namespace]):
            blob = torch.load(path, map_location="cpu", weights_only=True)
    except Exception as exc:
        if os.environ.get("PGIBBS_TRUST_CHECKPOINT") != "1":
            raise Exception("checkpoint %s needs more than tensors and argparse.Namespace to unpickle (%s); "
                            "set PGIBBS_TRUST_CHECKPOINT=1 if the file is trusted" % (path, exc))
        blob = torch.load(path, map_location="cpu", weights_only=False)
    ck_arch = getattr(blob.get("args", None), "arch", None) if isinstance(blob, dict) else None
    expected = {"roberta_large": "roberta_large", "esm1": "protein_bert_base", "msa_transformer": "msa_transformer"}
    if ck_arch is not None and arch in expected and ck_arch != expected[arch]:
        raise Exception("checkpoint %s holds a '%s' model, this class expects '%s'" % (path, ck_arch, expected[arch]))
    sd = blob["model"] if isinstance(blob, dict) and "model" in blob else blob
    return upgrade_state_dict(sd, arch)


PRECISION_LEVELS = {"fast": 0, "split_weights": 1, "split": 2}


def _precision(p):
    import os
    p = os.environ.get("PGIBBS_PRECISION", "fast") if p is None else p
    if p not in PRECISION_LEVELS:
        raise ValueError("precision must be one of %s, got %r" % (sorted(PRECISION_LEVELS), p))
    return p


class EngineModule:
    """Stands where the reference expects ``model.model``: ``eval()``, ``to(device)`` and
    ``__call__(tokens) -> {"logits": ...}`` (esm_sampler.py:62,80,223)."""

    def __init__(self, cfg, alphabet, state_dict, precision="fast"):
        self.cfg = cfg
        self.alphabet = alphabet
        self._state_dict = state_dict
        self.precision = precision
        self.engine = None
        self.device = "cpu"

    def eval(self):
        return self

    def to(self, device):
        device = str(device)
        if device == "cpu":
            return self  # weights stay on the host until a CUDA device is requested
        m = re.match(r"^cuda:([0-9]+)$", device)
        if not m:
            raise Exception("Invalid device: " + device)
        idx = int(m.group(1))
        if self.engine is None or self.engine.device_id != idx:
            if self.engine is not None:
                self.engine.close()
            self.engine = Engine(self.cfg, self.alphabet, idx, precision=self.precision)
            self.engine.load_state_dict(self._state_dict)
        self.device = device
        return self

    def cuda(self, index=0):
        return self.to("cuda:%d" % index)

    def require_engine(self):
        if self.engine is None:
            raise Exception("the B200 engine has no CPU path: construct the sampler with device='gpu' or 'cuda:N'")
        return self.engine

    def __call__(self, tokens, **kwargs):
        logits = self.require_engine().forward_logits(tokens.cpu() if isinstance(tokens, torch.Tensor) else tokens)
        return {"logits": logits, "representations": {}}


class _Triple:
    config_name = None
    alphabet_factory = staticmethod(Alphabet.esm1b)

    def __init__(self, state_dict=None, checkpoint=None, seed=0, precision=None, **cfg_overrides):
        """``precision``: "fast" (default; one pass of fp16 operands per GEMM) or "split" (every GEMM operand carried as
        an fp16 hi + lo pair, three tensor-core passes: logits within 1e-3 of fp32 per token row at full depth, about
        2.5x the step time -- DESIGN.md section 3).  ``None`` reads ``PGIBBS_PRECISION`` (default "fast")."""
        self.cfg = get_config(self.config_name, **cfg_overrides)
        self.alphabet = self.alphabet_factory()
        if checkpoint is not None:
            state_dict = load_checkpoint(checkpoint, self.cfg["arch"])
        if state_dict is None:
            state_dict = synthetic_state_dict(self.cfg, seed)
        self.model = EngineModule(self.cfg, self.alphabet, state_dict, precision=_precision(precision))
        self.batch_converter = self.alphabet.get_batch_converter()


class ESM1b(_Triple):
    config_name = "esm1b_t33_650M_UR50S"


class ESM1v(_Triple):
    config_name = "esm1v_t33_650M_UR90S_1"


class ESM6(_Triple):
    config_name = "esm1_t6_43M_UR50S"
    alphabet_factory = staticmethod(Alphabet.esm1)


class ESM12(_Triple):
    config_name = "esm1_t12_85M_UR50S"
    alphabet_factory = staticmethod(Alphabet.esm1)


class ESM34(_Triple):
    config_name = "esm1_t34_670M_UR50S"
    alphabet_factory = staticmethod(Alphabet.esm1)


class ESM2_t6_8M(_Triple):
    config_name = "esm2_t6_8M_UR50D"


class ESM2_t30_150M(_Triple):
    config_name = "esm2_t30_150M_UR50D"


class ESM2_t33_650M(_Triple):
    config_name = "esm2_t33_650M_UR50D"


class ESM_MSA1(_Triple):
    config_name = "esm_msa1b_t12_100M_UR50S"
    alphabet_factory = staticmethod(Alphabet.msa)


class CustomModel(_Triple):
    """Arbitrary geometry (tests): ``CustomModel(cfg_dict, seed=...)``."""

    def __init__(self, cfg, state_dict=None, seed=0, precision=None):
        self.cfg = dict(cfg)
        self.alphabet = {"msa_transformer": Alphabet.msa, "esm1": Alphabet.esm1}.get(cfg["arch"], Alphabet.esm1b)()
        if state_dict is None:
            state_dict = synthetic_state_dict(self.cfg, seed)
        self.model = EngineModule(self.cfg, self.alphabet, state_dict, precision=_precision(precision))
        self.batch_converter = self.alphabet.get_batch_converter()
