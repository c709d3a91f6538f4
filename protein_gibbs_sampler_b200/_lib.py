"""ctypes binding of libpgibbs.so (the C ABI declared in include/pgibbs.h).

The shared library is the product's only compute path.  If it is missing or cannot be loaded this module
raises -- there is no CPU or PyTorch fallback.
"""
import ctypes
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
# (PGIBBS_LIB_PATH: A/B tooling only -- e.g. profiling a previous build of the same ABI)
LIB_PATH = os.environ.get("PGIBBS_LIB_PATH") or os.path.join(_HERE, "libpgibbs.so")

c_i32, c_i64, c_f32 = ctypes.c_int32, ctypes.c_int64, ctypes.c_float
c_void_p, c_char_p = ctypes.c_void_p, ctypes.c_char_p
P = ctypes.POINTER


class ModelConfig(ctypes.Structure):
    _fields_ = [(n, c_i32) for n in (
        "arch", "layers", "embed_dim", "heads", "ffn_dim", "vocab", "max_positions",
        "token_dropout", "padding_idx", "mask_idx", "cls_idx", "eos_idx")]


# name -> (restype, argtypes); must list every symbol include/pgibbs.h declares
SIGNATURES = {
    "pgibbs_last_error": (c_char_p, []),
    "pgibbs_version": (c_char_p, []),
    "pgibbs_create": (c_i32, [P(ModelConfig), c_i32, P(c_void_p)]),
    "pgibbs_destroy": (c_i32, [c_void_p]),
    "pgibbs_set_stream": (c_i32, [c_void_p, c_void_p, c_i32]),
    "pgibbs_load_weight": (c_i32, [c_void_p, c_char_p, c_void_p, c_i64]),
    "pgibbs_set_precision": (c_i32, [c_void_p, c_i32]),
    "pgibbs_finalize_weights": (c_i32, [c_void_p]),
    "pgibbs_set_tokens": (c_i32, [c_void_p, c_void_p, c_i32, c_i32, c_i32]),
    "pgibbs_get_tokens": (c_i32, [c_void_p, c_void_p]),
    "pgibbs_set_schedule": (c_i32, [c_void_p, c_void_p, c_i64, c_i32, c_i32, c_i64, c_i64, c_i32]),
    "pgibbs_set_noise": (c_i32, [c_void_p, c_void_p, c_i64, c_i32]),
    "pgibbs_set_device_rng": (c_i32, [c_void_p, ctypes.c_uint64]),
    "pgibbs_set_chain_offset": (c_i32, [c_void_p, c_i64]),
    "pgibbs_run": (c_i32, [c_void_p, c_i32, c_i32, c_i64, c_i32, c_f32, c_i32, c_void_p, c_i32]),
    "pgibbs_run_single": (c_i32, [c_void_p, c_i32, c_i32, c_i64, c_i32, c_f32, c_i32, c_i32, c_void_p, c_i32]),
    "pgibbs_forward_logits": (c_i32, [c_void_p, c_void_p, c_i32, c_i32, c_i32, c_void_p]),
    "pgibbs_score": (c_i32, [c_void_p, c_void_p, c_i32, c_i32, c_void_p]),
    "pgibbs_sync": (c_i32, [c_void_p]),
    "pgibbs_debug_read": (c_i32, [c_void_p, c_char_p, c_void_p, c_i64]),
    "pgibbs_debug_tail_plan": (c_i32, [c_i32, c_i32, c_i32, c_i32, c_i32, c_i32, P(c_i32)]),
    "pgibbs_debug_layer_limit": (c_i32, [c_void_p, c_i32]),
    "pgibbs_profile_enable": (c_i32, [c_void_p, c_i32]),
    "pgibbs_profile_read": (c_i32, [c_void_p, c_void_p, c_void_p, c_void_p, c_i32, P(c_i32)]),
    "pgibbs_launch_count": (c_i64, [c_void_p]),
    "pgibbs_debug_attention_trace": (c_i32, [c_void_p, c_i32]),
    "pgibbs_op_gemm": (c_i32, [c_i32, c_void_p, c_void_p, c_void_p, c_void_p, c_i32, c_i32, c_i32, c_i32, c_i32, c_i32,
                               P(c_f32), c_i32]),
    "pgibbs_op_attention": (c_i32, [c_i32, c_void_p, c_void_p, c_i32, c_i32, c_i32, c_i32, P(c_f32), c_i32]),
    "pgibbs_op_sample": (c_i32, [c_i32, c_void_p, c_void_p, c_i32, c_i32, c_void_p, c_i32, c_i32, c_f32, c_void_p]),
}

_lib = None


class EngineError(Exception):
    pass


def load():
    """Load libpgibbs.so once; raise loudly if the CUDA extension has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise EngineError(
            "libpgibbs.so not found at %s: build the CUDA extension first "
            "(python -c 'import __graft_entry__ as g; g.build()'). There is no CPU fallback." % LIB_PATH)
    lib = ctypes.CDLL(LIB_PATH)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)  # AttributeError here = header/library mismatch
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib


def check(rc):
    if rc != 0:
        raise EngineError(load().pgibbs_last_error().decode("utf-8", "replace"))
