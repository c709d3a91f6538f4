"""Python handle on one pgibbs engine (one per GPU).  Thin: every method is one C-ABI call."""
import ctypes
import math

import numpy as np
import torch

from . import _lib
from ._lib import EngineError, check

_ARCH = {"roberta_large": 0, "esm2": 1, "msa_transformer": 2, "esm1": 3}
INT64_MAX = (1 << 63) - 1


def _ptr(t):
    """Address of a contiguous torch tensor / numpy array (host or device)."""
    if isinstance(t, torch.Tensor):
        assert t.is_contiguous()
        return ctypes.c_void_p(t.data_ptr())
    assert t.flags["C_CONTIGUOUS"]
    return ctypes.c_void_p(t.ctypes.data)


class Engine:
    def __init__(self, cfg, alphabet, device_id=0, precision="fast"):
        self.lib = _lib.load()
        self.cfg = dict(cfg)
        mc = _lib.ModelConfig(
            arch=_ARCH[cfg["arch"]], layers=cfg["layers"], embed_dim=cfg["embed_dim"], heads=cfg["heads"],
            ffn_dim=cfg["ffn_dim"], vocab=cfg["vocab"], max_positions=cfg["max_positions"],
            token_dropout=int(cfg["token_dropout"]), padding_idx=alphabet.padding_idx, mask_idx=alphabet.mask_idx,
            cls_idx=alphabet.cls_idx, eos_idx=alphabet.eos_idx)
        h = ctypes.c_void_p()
        check(self.lib.pgibbs_create(ctypes.byref(mc), int(device_id), ctypes.byref(h)))
        self.h = h
        self.device_id = int(device_id)
        self.shape = None
        self.precision = precision
        level = {"fast": 0, "split_weights": 1, "split": 2}[precision]
        if level:
            check(self.lib.pgibbs_set_precision(self.h, level))

    def close(self):
        if getattr(self, "h", None):
            self.lib.pgibbs_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # ---- weights
    def load_state_dict(self, sd):
        """fp32 tensors keyed like a fair-esm state dict (host or same-device CUDA tensors)."""
        for name, t in sd.items():
            if not isinstance(t, torch.Tensor) or not t.is_floating_point():
                continue
            if name.startswith("contact_head") or name.endswith("inv_freq") or name.endswith("_float_tensor"):
                continue
            if name == "lm_head.weight":
                continue  # tied to embed_tokens.weight (RobertaLMHead)
            t = t.detach().to(torch.float32).contiguous()
            check(self.lib.pgibbs_load_weight(self.h, name.encode(), _ptr(t), t.numel()))
        if self.cfg["arch"] == "esm1":
            # sinusoidal positions are a fixed table: computed here exactly as fair-esm does (float32) and handed to
            # the engine through the learned-position slot (rows 2.. are positions 0.., one spare row for the
            # internal bias-key/value slot)
            from .weights import sinusoidal_table
            tbl = sinusoidal_table(self.cfg["max_positions"] + 3, self.cfg["embed_dim"]).contiguous()
            check(self.lib.pgibbs_load_weight(self.h, b"embed_positions.weight", _ptr(tbl), tbl.numel()))
        check(self.lib.pgibbs_finalize_weights(self.h))

    # ---- state
    def set_stream(self, cuda_stream_ptr):
        """Launch on an external CUDA stream (e.g. ``torch.cuda.current_stream().cuda_stream``; 0 is the CUDA
        default stream).  ``None`` returns to the engine's own stream."""
        if cuda_stream_ptr is None:
            check(self.lib.pgibbs_set_stream(self.h, None, 0))
        else:
            check(self.lib.pgibbs_set_stream(self.h, ctypes.c_void_p(cuda_stream_ptr), 1))

    def set_tokens(self, tokens):
        t = torch.as_tensor(tokens)
        if t.dim() == 2:
            t = t[:, None, :]
        t = t.to(torch.int32).contiguous()
        B, R, T = t.shape
        check(self.lib.pgibbs_set_tokens(self.h, _ptr(t), B, R, T))
        self.shape = (B, R, T)

    def get_tokens(self):
        B, R, T = self.shape
        out = torch.empty((B, R, T), dtype=torch.int32)
        check(self.lib.pgibbs_get_tokens(self.h, _ptr(out)))
        return out.to(torch.int64)

    def set_schedule(self, positions, n_iters, P, iter_stride, chain_stride, has_duplicates=False):
        p = np.ascontiguousarray(positions, dtype=np.int32)
        check(self.lib.pgibbs_set_schedule(self.h, _ptr(p), p.size, n_iters, P, iter_stride, chain_stride,
                                           int(has_duplicates)))

    def set_noise(self, noise, stride=0):
        if noise is None:
            check(self.lib.pgibbs_set_noise(self.h, None, 0, 0))
        else:
            n = torch.as_tensor(noise, dtype=torch.float32).contiguous()
            check(self.lib.pgibbs_set_noise(self.h, _ptr(n), n.numel(), int(stride)))

    def set_device_rng(self, seed):
        check(self.lib.pgibbs_set_device_rng(self.h, ctypes.c_uint64(seed & ((1 << 64) - 1))))

    def set_chain_offset(self, first_chain):
        check(self.lib.pgibbs_set_chain_offset(self.h, int(first_chain)))

    @staticmethod
    def _burnin(b):
        """`ii < burnin` for integer ii  <=>  ii < ceil(burnin); inf -> never leave burn-in."""
        if b == float("inf"):
            return INT64_MAX
        if b != b:  # nan compares False
            return 0
        return int(min(max(math.ceil(b), 0), INT64_MAX))

    @staticmethod
    def _temperature(t):
        """None -> NaN (the ABI's "no temperature"); any other value divides the logits as the reference's
        generate_step does (esm_sampler.py:24-25), including negative ones.  0 would make every logit +-inf / nan,
        which torch's Categorical rejects in the reference; it is rejected here too."""
        if t is None:
            return float("nan")
        t = float(t)
        if t == 0.0 or t != t:
            raise ValueError("temperature must be a non-zero number or None, got %r" % (t,))
        return t

    def run(self, first_iter, num_iters, burnin, top_k, temperature, mask, valid_ids):
        v = np.ascontiguousarray(valid_ids, dtype=np.int32)
        temp = self._temperature(temperature)
        check(self.lib.pgibbs_run(self.h, first_iter, num_iters, self._burnin(burnin), int(top_k), temp, int(bool(mask)),
                                  _ptr(v), v.size))

    def run_single(self, first_iter, num_iters, burnin, top_k, temperature, mask_row, target_row, valid_ids):
        v = np.ascontiguousarray(valid_ids, dtype=np.int32)
        temp = self._temperature(temperature)
        check(self.lib.pgibbs_run_single(self.h, first_iter, num_iters, self._burnin(burnin), int(top_k), temp,
                                         int(mask_row), int(target_row), _ptr(v), v.size))

    def forward_logits(self, tokens):
        t = torch.as_tensor(tokens)
        squeeze = t.dim() == 2
        if squeeze:
            t = t[:, None, :]
        t = t.to(torch.int32).contiguous()
        B, R, T = t.shape
        out = torch.empty((B, R, T, self.cfg["vocab"]), dtype=torch.float32)
        check(self.lib.pgibbs_forward_logits(self.h, _ptr(t), B, R, T, _ptr(out)))
        self.shape = (B, R, T)
        return out[:, 0] if squeeze else out

    def score(self, targets, mask=True, row=-1):
        """log_softmax(logits)[target] for every scheduled slot of the resident tokens (pgibbs_score); targets is
        [n_chains, P] int, < 0 = padding slot.  Returns float32 of the same shape."""
        t = torch.as_tensor(targets).to(torch.int32).contiguous()
        out = torch.empty(t.shape, dtype=torch.float32)
        check(self.lib.pgibbs_score(self.h, _ptr(t), int(bool(mask)), int(row), _ptr(out)))
        return out

    def sync(self):
        check(self.lib.pgibbs_sync(self.h))

    # ---- debug / profiling
    def debug_read(self, which, numel):
        out = torch.empty(numel, dtype=torch.float32)
        check(self.lib.pgibbs_debug_read(self.h, which.encode(), _ptr(out), numel))
        return out

    def debug_layer_limit(self, n):
        check(self.lib.pgibbs_debug_layer_limit(self.h, int(n)))

    def profile_enable(self, on=True):
        check(self.lib.pgibbs_profile_enable(self.h, int(on)))

    def profile_read(self):
        cap = 32
        names = (ctypes.c_char * 32 * cap)()
        ms = (ctypes.c_float * cap)()
        cnt = (ctypes.c_int32 * cap)()
        n = ctypes.c_int32(0)
        check(self.lib.pgibbs_profile_read(self.h, names, ms, cnt, cap, ctypes.byref(n)))
        return {names[i].value.decode(): (float(ms[i]), int(cnt[i])) for i in range(n.value)}

    def launch_count(self):
        return int(self.lib.pgibbs_launch_count(self.h))


# ---- stand-alone operators (tests)
def op_gemm(A, B, bias=None, C=None, epilogue=5, block_n=0, device_id=0, reps=0, cta_group=0):
    lib = _lib.load()
    A = torch.as_tensor(A, dtype=torch.float32).contiguous()
    B = torch.as_tensor(B, dtype=torch.float32).contiguous()
    M, K = A.shape
    N = B.shape[0]
    out = torch.zeros((M, N), dtype=torch.float32) if C is None else C.clone().contiguous()
    b = None if bias is None else torch.as_tensor(bias, dtype=torch.float32).contiguous()
    ms = ctypes.c_float(0)
    check(lib.pgibbs_op_gemm(device_id, _ptr(A), _ptr(B), None if b is None else _ptr(b), _ptr(out), M, N, K,
                             epilogue, block_n, cta_group, ctypes.byref(ms), reps))
    return (out, ms.value) if reps else out


def op_attention(qkv, n_seq, T, heads, head_dim, device_id=0, reps=0):
    lib = _lib.load()
    q = torch.as_tensor(qkv, dtype=torch.float32).contiguous()
    out = torch.empty((n_seq * T, heads * head_dim), dtype=torch.float32)
    ms = ctypes.c_float(0)
    check(lib.pgibbs_op_attention(device_id, _ptr(q), _ptr(out), n_seq, T, heads, head_dim, ctypes.byref(ms), reps))
    return (out, ms.value) if reps else out


def op_sample(logits, noise, valid_ids, top_k=0, temperature=None, device_id=0):
    lib = _lib.load()
    l = torch.as_tensor(logits, dtype=torch.float32).contiguous()
    rows, V = l.shape
    v = np.ascontiguousarray(valid_ids, dtype=np.int32)
    n = None if noise is None else torch.as_tensor(noise, dtype=torch.float32).contiguous()
    out = torch.empty(rows, dtype=torch.int32)
    check(lib.pgibbs_op_sample(device_id, _ptr(l), None if n is None else _ptr(n), rows, V, _ptr(v), v.size, int(top_k),
                               Engine._temperature(temperature), _ptr(out)))
    return out.to(torch.int64)
