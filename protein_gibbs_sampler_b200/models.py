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


def upgrade_state_dict(sd):
    """Strip the prefixes fair-esm checkpoints carry (esm.pretrained's key rewriting)."""
    out = {}
    for k, v in sd.items():
        k = re.sub(r"^(encoder\.)?sentence_encoder\.", "", k)
        k = re.sub(r"^encoder\.", "", k)
        k = re.sub(r"^msa\.", "", k)
        out[k] = v
    return out


class EngineModule:
    """Stands where the reference expects ``model.model``: ``eval()``, ``to(device)`` and
    ``__call__(tokens) -> {"logits": ...}`` (esm_sampler.py:62,80,223)."""

    def __init__(self, cfg, alphabet, state_dict):
        self.cfg = cfg
        self.alphabet = alphabet
        self._state_dict = state_dict
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
            self.engine = Engine(self.cfg, self.alphabet, idx)
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

    def __init__(self, state_dict=None, checkpoint=None, seed=0, **cfg_overrides):
        self.cfg = get_config(self.config_name, **cfg_overrides)
        self.alphabet = self.alphabet_factory()
        if checkpoint is not None:
            blob = torch.load(checkpoint, map_location="cpu", weights_only=False)
            state_dict = upgrade_state_dict(blob["model"] if "model" in blob else blob)
        if state_dict is None:
            state_dict = synthetic_state_dict(self.cfg, seed)
        self.model = EngineModule(self.cfg, self.alphabet, state_dict)
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

    def __init__(self, cfg, state_dict=None, seed=0):
        self.cfg = dict(cfg)
        self.alphabet = {"msa_transformer": Alphabet.msa, "esm1": Alphabet.esm1}.get(cfg["arch"], Alphabet.esm1b)()
        if state_dict is None:
            state_dict = synthetic_state_dict(self.cfg, seed)
        self.model = EngineModule(self.cfg, self.alphabet, state_dict)
        self.batch_converter = self.alphabet.get_batch_converter()
