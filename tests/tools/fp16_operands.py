"""TEST INFRASTRUCTURE (not a product path): the fp32 CPU oracle with exactly the tensors the CUDA kernels round
to fp16 rounded at the same points -- GEMM inputs and weights (`tcgen05.mma kind::f16` operands), q / k after
scale + RoPE, v, the softmax probabilities and the attention output.  Accumulation stays fp32 as in TMEM.

Purpose: separate "the kernels compute something else" from "fp16 operands round".  The engine must agree with
THIS model far more tightly (a few 1e-4: accumulation order, the attention kernel rounding un-normalised
probabilities) than with the pure fp32 oracle (up to 1e-3); `tests/tools/stage_parity.py` prints both, stage by stage.
Note that the two never agree much better than each does with fp32: a difference of a few percent of an fp16 ulp
upstream is enough to flip the rounding of a comparable share of the elements downstream.
"""
import torch


def _r(t):
    return t.half().float()


def with_fp16_operands(oracle_model):
    """Patch an `OracleModel` wrapping an `ESMOracle` in place and return it."""
    mod = oracle_model.model

    def lin(x, prefix):
        return torch.matmul(_r(x), _r(mod.sd[prefix + ".weight"]).t()) + mod.sd[prefix + ".bias"]

    mod._lin = lin
    mod.mm = lambda a, b: torch.matmul(_r(a), _r(b))
    return oracle_model
