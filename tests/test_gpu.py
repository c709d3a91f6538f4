"""Parity tests proper: the CUDA path (through the C ABI) against the CPU oracle on the same seeded inputs,
against the golden vectors produced by the reference, and size-independent properties at full size.

Tolerances: integer / index / token work is bit-exact.  Floating point: logits within 1e-3 of the fp32 oracle,
measured as max|delta| / max|logit| over the batch (BASELINE.json north_star: "MLM logits to <= 1e-3 relative");
operator-level tolerances are stated per test.
"""
import random

import pytest
import torch

pytestmark = pytest.mark.gpu

LOGIT_TOL = 1e-3


def rel(got, want):
    return ((got - want).abs().max() / want.abs().max()).item()


def rel_rows(got, want, valid_ids=None):
    """SURVEY 8c's stricter metric: every token row on its own, max|delta| / max|logit| of THAT row (the batch-wide
    maximum in `rel` is dominated by the rows with the largest logits); with `valid_ids`, numerator and denominator
    run over the candidate residues only (what generate_step reads).  Returns the worst row."""
    g, w = got.reshape(-1, got.shape[-1]), want.reshape(-1, want.shape[-1])
    if valid_ids is not None:
        g, w = g[:, valid_ids], w[:, valid_ids]
    return ((g - w).abs().amax(-1) / w.abs().amax(-1)).max().item()


# Documented bound of the single-pass fp16-operand ("fast") mode on the per-row metric at full depth (33 layers):
# CPU emulation of the operand rounding gives 1.4-2.2e-3 over weight / token seeds (tests/tools/precision_study3.py).
ROW_TOL_FAST = 2.5e-3
# ... and on the batch-wide metric: 0.6-1.0e-3 (ESM-1b), 0.9-1.6e-3 (ESM-2 650M) in the same emulation; measured on
# the B200 at the BASELINE shapes 1.0e-3 (config 2) and 1.0-1.3e-3 (config 4).  Split-operand mode asserts 1e-3.
BATCH_TOL_FAST_DEEP = 1.7e-3


@pytest.fixture(scope="module", autouse=True)
def _need_gpu(gpu_lib):
    assert torch.cuda.is_available(), "GPU tests need a CUDA device"


def make(cfg, seed, rng="replay", precision="fast"):
    from protein_gibbs_sampler_b200 import models
    from protein_gibbs_sampler_b200.esm_msa_sampler import ESM_MSA_sampler
    from protein_gibbs_sampler_b200.esm_sampler import ESM_sampler
    from protein_gibbs_sampler_b200.weights import synthetic_state_dict
    sd = synthetic_state_dict(cfg, seed)
    m = models.CustomModel(cfg, state_dict=sd, precision=precision)
    cls = ESM_MSA_sampler if cfg["arch"] == "msa_transformer" else ESM_sampler
    return cls(m, device="cuda:0", rng=rng), sd


# ------------------------------------------------------------------------------------------- operators


def test_gemm_last_wave_k_split_opt_in(monkeypatch):
    """PGIBBS_GEMM_SPLIT=1 (off by default: DESIGN.md section 4): FC2's partly filled last wave cut along K, the parts'
    reduce-adds ordered by flags.  Same value as the unsplit kernel up to the fp32 regrouping x + p0 + p1, and
    bit-identical from run to run."""
    from protein_gibbs_sampler_b200.engine import op_gemm
    torch.manual_seed(0)
    M, N, K = 16512, 1280, 5120          # 325 pair tiles on 74 pairs: 29 tiles in the last wave
    A, B, bias, C0 = torch.randn(M, K) * 0.5, torch.randn(N, K) * 0.05, torch.randn(N), torch.randn(M, N)
    monkeypatch.setenv("PGIBBS_GEMM_SPLIT", "0")
    plain = op_gemm(A, B, bias, C=C0, epilogue=2)
    monkeypatch.setenv("PGIBBS_GEMM_SPLIT", "1")
    split1 = op_gemm(A, B, bias, C=C0, epilogue=2)
    split2 = op_gemm(A, B, bias, C=C0, epilogue=2)
    monkeypatch.setenv("PGIBBS_GEMM_SPLIT", "0")
    op_gemm(A[:128], B, bias, C=C0[:128], epilogue=2)     # leaves the process-wide switch off again
    assert torch.equal(split1, split2)
    assert not torch.equal(split1, plain)                 # the split did happen ...
    assert rel(split1, plain) < 1e-5                      # ... and only regroups an fp32 sum (operand rounding: 3e-4)

@pytest.mark.parametrize("M,N,K,bn", [(128, 64, 64, 64), (300, 320, 320, 64), (516, 960, 320, 192),
                                      (1000, 1280, 1280, 256), (129, 336, 128, 128), (2050, 5120, 1280, 256)])
def test_gemm_operator(M, N, K, bn):
    from protein_gibbs_sampler_b200.engine import op_gemm
    g = torch.Generator().manual_seed(M + N)
    A, B, bias = torch.randn(M, K, generator=g) * 0.5, torch.randn(N, K, generator=g) * 0.5, torch.randn(N, generator=g)
    want = A.half().float() @ B.half().float().t() + bias        # fp16 operands, fp32 accumulate
    assert rel(op_gemm(A, B, bias, epilogue=5, block_n=bn), want) < 2e-5
    assert rel(op_gemm(A, B, bias, epilogue=0, block_n=bn), want) < 1e-3          # fp16 output rounding
    assert rel(op_gemm(A, B, bias, epilogue=4, block_n=bn), torch.nn.functional.gelu(want)) < 2e-5
    C0 = torch.randn(M, N, generator=g)
    assert rel(op_gemm(A, B, bias, C=C0, epilogue=2, block_n=bn), want + C0) < 2e-5


@pytest.mark.parametrize("M,N,K,bn", [(300, 320, 320, 64), (129, 336, 128, 128), (2050, 5120, 1280, 256), (16512, 1280, 1280, 256)])
def test_gemm_output_paths_bit_identical(M, N, K, bn, monkeypatch):
    """The plain-store epilogues write either straight from registers (st.global.v8, the default) or through
    shared-memory staging and the TMA engine (PGIBBS_EPI_DIRECT=0): same values, same clipping at the M / N edges (the
    TMA path gets it from the tensor map, the direct path from its own guards) -- every output bit must agree."""
    from protein_gibbs_sampler_b200.engine import op_gemm
    g = torch.Generator().manual_seed(M ^ N)
    A, B, bias = torch.randn(M, K, generator=g) * 0.5, torch.randn(N, K, generator=g) * 0.5, torch.randn(N, generator=g)
    for epi in (0, 1, 4, 5):
        monkeypatch.setenv("PGIBBS_EPI_DIRECT", "6")
        direct = op_gemm(A, B, bias, epilogue=epi, block_n=bn)
        monkeypatch.setenv("PGIBBS_EPI_DIRECT", "0")
        staged = op_gemm(A, B, bias, epilogue=epi, block_n=bn)
        assert torch.equal(direct, staged), "epilogue %d" % epi
    monkeypatch.setenv("PGIBBS_EPI_DIRECT", "6")
    op_gemm(A[:128], B, bias, epilogue=0, block_n=bn)     # leaves the process-wide switch at its default (direct) again


def test_gemm_operator_config2_rows_vs_torch():
    """The GEMM shapes of BASELINE config 2 (M = 64 x 258 = 16512 token rows: 65 row blocks, partly filled last wave)
    against torch on the same fp16-rounded operands: out-projection (residual epilogue) and FC1 (GELU epilogue)."""
    from protein_gibbs_sampler_b200.engine import op_gemm
    g = torch.Generator().manual_seed(16512)
    M = 16512
    for N, K, epi in [(1280, 1280, 2), (5120, 1280, 4), (1280, 5120, 2)]:
        A, B, bias = torch.randn(M, K, generator=g) * 0.5, torch.randn(N, K, generator=g) * 0.05, torch.randn(N, generator=g)
        want = A.half().float() @ B.half().float().t() + bias
        if epi == 2:
            C0 = torch.randn(M, N, generator=g)
            got = op_gemm(A, B, bias, C=C0, epilogue=2)
            want = want + C0
        else:
            got = op_gemm(A, B, bias, epilogue=4)
            want = torch.nn.functional.gelu(want)
        assert rel(got, want) < 2e-5, (N, K, epi, rel(got, want))
        # every row block was written (the last, partly filled 256-row tile included)
        assert rel(got[-300:], want[-300:]) < 2e-5 and rel(got[:300], want[:300]) < 2e-5


@pytest.mark.parametrize("n_seq,T,H,Dh", [(2, 64, 2, 64), (2, 27, 20, 16), (3, 258, 4, 64), (2, 130, 3, 32),
                                          (1, 1024, 2, 64), (3, 5, 2, 64)])
def test_attention_operator(n_seq, T, H, Dh):
    from protein_gibbs_sampler_b200.engine import op_attention
    qkv = torch.randn(n_seq * T, 3 * H * Dh, generator=torch.Generator().manual_seed(T)) * 0.7
    x = qkv.half().float().view(n_seq, T, 3, H, Dh)
    q, k, v = (x[:, :, i].transpose(1, 2) for i in range(3))
    want = (torch.softmax(q @ k.transpose(-1, -2), -1) @ v).transpose(1, 2).reshape(n_seq * T, H * Dh)
    assert rel(op_attention(qkv, n_seq, T, H, Dh), want) < 3e-3   # P and ctx are rounded to fp16


@pytest.mark.parametrize("T,hot_keys", [(258, (40,)), (258, (100, 250)), (300, (5, 70, 130, 257, 299)), (1024, (33, 200, 1000)),
                                        (514, (513,)), (200, (96, 97, 191))])
def test_attention_operator_growing_maximum(T, hot_keys):
    """Keys whose scores exceed everything before them by far more than the lazy-rescale threshold (2^11), placed in the
    first and in the second half of 64-key sub-blocks, early and late in the sequence: exercises the O rescale, the
    rescale of the half sub-block of P that is already in TMEM, and the trailing-row path of the tcgen05 attention."""
    from protein_gibbs_sampler_b200.engine import op_attention
    n_seq, H, Dh = 2, 3, 64
    g = torch.Generator().manual_seed(T + len(hot_keys))
    x = torch.randn(n_seq, T, 3, H, Dh, generator=g) * 0.3
    x[:, :, 2] *= 3
    u = torch.nn.functional.normalize(torch.randn(H, Dh, generator=g), dim=-1)
    x[:, :, 0] += 3.0 * u                      # every query has a component along u ...
    for j, key in enumerate(hot_keys):         # ... and the hot keys are increasingly aligned with it
        x[:, key, 1] += (8.0 + 6.0 * j) * u
    qkv = x.reshape(n_seq * T, 3 * H * Dh)
    xh = qkv.half().float().view(n_seq, T, 3, H, Dh)
    q, k, v = (xh[:, :, i].transpose(1, 2) for i in range(3))
    s = q @ k.transpose(-1, -2)
    assert float((s.max(-1).values - s[..., :hot_keys[0]].max(-1).values).min()) > 9.0   # > 2^11 in every row
    want = (torch.softmax(s, -1) @ v).transpose(1, 2).reshape(n_seq * T, H * Dh)
    assert rel(op_attention(qkv, n_seq, T, H, Dh), want) < 3e-3


def test_sampler_tail_bit_exact_vs_oracle():
    from oracle.sampler_tail import generate_step_with_noise
    from protein_gibbs_sampler_b200.engine import op_sample
    g = torch.Generator().manual_seed(0)
    for valid, top_k, temp in [(list(range(4, 24)), 0, None), (list(range(4, 24)), 3, None),
                               (list(range(4, 24)) + [30], 5, 0.7), ([3, 5, 1], 2, None), (list(range(4, 24)), 1, 2.0),
                               (list(range(4, 24)), 50, None), (list(range(4, 24)), 4, -1.5)]:
        rows = 3000
        logits = torch.randn(rows, 33, generator=g) * 2
        noise = torch.empty(rows, len(valid)).exponential_(1, generator=g)
        got = op_sample(logits, noise, valid, top_k=top_k, temperature=temp)
        want = torch.tensor([generate_step_with_noise(logits[i], noise[i], valid, top_k, temp) for i in range(rows)])
        assert torch.equal(got, want)


def test_generate_step_golden_and_statistics(golden):
    """Reference's own draws under a seeded torch RNG, then its statistical tests (test_esm_sampler.py:185-253)."""
    from protein_gibbs_sampler_b200.esm_sampler import generate_step
    for c in golden["generate_step"]:
        torch.manual_seed(c["torch_seed"])
        got = int(generate_step(torch.tensor(c["logits"]), c["gen_idx"], temperature=c["temperature"],
                                top_k=c["top_k"], sample=c["sample"], valid_idx=c["valid_idx"]))
        assert got == c["token"]
    torch.manual_seed(0)
    out = torch.ones(1, 6)
    counts = [0] * 6
    for _ in range(1000):
        counts[int(generate_step(out, 0))] += 1
    assert all(c > 100 for c in counts)
    counts = [0] * 6
    for _ in range(1000):
        counts[int(generate_step(out, 0, valid_idx=[1, 3, 5]))] += 1
    assert counts[0] == counts[2] == counts[4] == 0 and all(counts[i] > 200 for i in (1, 3, 5))
    out = torch.tensor([[0.4, 0.2, 0.4, 0.2, 0.1, 0.1]])
    for valid in ([1, 3, 5], [3, 5, 1]):
        counts = [0] * 6
        for _ in range(1000):
            counts[int(generate_step(out, 0, top_k=2, valid_idx=valid))] += 1
        assert counts[1] > 400 and counts[3] > 400 and counts[5] == 0


# --------------------------------------------------------------------------------------------- forward
def _tokens(cfg, shape, seed):
    g = torch.Generator().manual_seed(seed)
    tok = torch.randint(4, 24, shape, generator=g)
    esm1 = cfg["arch"] == "esm1"
    cls, mask = (32, 33) if esm1 else (0, 32)
    tok[..., 0] = cls
    if cfg["arch"] not in ("msa_transformer", "esm1"):
        tok[..., -1] = 2
    flat = tok.view(-1, shape[-1])
    flat[0, 3:9] = mask
    if flat.shape[0] > 1:
        flat[1, 1:shape[-1] - 1] = mask     # a fully masked chain: largest token-dropout rescale
    return tok


FORWARD_CASES = [
    ("esm2", 2, 128, 2, 256, (2, 24)), ("roberta_large", 2, 128, 2, 256, (2, 24)),
    ("esm2", 6, 320, 20, 1280, (2, 27)),                    # BASELINE config 1 geometry (esm2_t6_8M)
    ("roberta_large", 3, 256, 4, 512, (3, 130)), ("esm2", 2, 640, 20, 2560, (2, 70)),
    ("roberta_large", 33, 1280, 20, 5120, (2, 66)),         # full-depth ESM-1b (config 2/5 model)
    ("esm2", 33, 1280, 20, 5120, (2, 40)),                  # full-depth ESM-2 650M (config 4 model)
    ("esm1", 2, 128, 2, 256, (2, 24)), ("esm1", 2, 128, 8, 256, (3, 131)),  # ESM-1: head_dim 64 / 16, bias key/value
    ("esm1", 6, 768, 12, 3072, (2, 66)),                    # esm1_t6_43M geometry (the reference's esm6)
    ("esm1", 2, 256, 4, 512, (2, 257)),                     # 257 tokens + the bias slot = 258: tcgen05 attention tail
    ("msa_transformer", 2, 128, 2, 256, (2, 4, 17)), ("msa_transformer", 2, 768, 12, 3072, (1, 8, 33)),
    ("msa_transformer", 2, 128, 2, 256, (2, 1, 9)),         # single-row MSA (pgen_msa_revised --alignment_size 1)
    ("msa_transformer", 2, 128, 2, 256, (1, 3, 200)),       # tcgen05 tied row attention, two query tiles
    ("msa_transformer", 2, 128, 2, 256, (1, 5, 129)),       # ... second tile with a single valid row (L = 128)
    ("msa_transformer", 2, 128, 2, 256, (1, 2, 300)),       # wider than 256 columns: mma.sync row attention
    ("msa_transformer", 2, 128, 4, 256, (1, 3, 70)),        # head_dim 32: mma.sync kernels throughout
    ("msa_transformer", 2, 128, 2, 256, (1, 40, 12)),       # more than 32 rows: generic column-attention kernel
    ("msa_transformer", 2, 192, 3, 256, (1, 20, 19)),       # odd head count: generic column-attention kernel
    ("msa_transformer", 12, 768, 12, 3072, (1, 6, 40)),     # full-depth MSA-1b (config 3 model)
    # ---- full depth AT the BASELINE token shapes (every perf number is quoted on these)
    ("roberta_large", 33, 1280, 20, 5120, (2, 258)),        # config 2: L = 256
    ("roberta_large", 33, 1280, 20, 5120, (1, 1024)),       # config 5: L = 1022, the longest the positions allow
    ("esm2", 33, 1280, 20, 5120, (2, 514)),                 # config 4: L = 512
    ("msa_transformer", 12, 768, 12, 3072, (1, 32, 129)),   # config 3: 32 rows x L = 128
]


def _forward_errors(arch, layers, d, H, F, shape, precision="fast", weights_seed=3):
    from oracle.fair_esm import OracleModel
    from protein_gibbs_sampler_b200.config import tiny_config
    cfg = tiny_config(arch, layers, d, H, F)
    s, sd = make(cfg, weights_seed, precision=precision)
    tok = _tokens(cfg, shape, 5)
    got = s.model.model(tok)["logits"]
    want = OracleModel(cfg, sd).model(tok)["logits"]
    assert got.shape == want.shape and bool(torch.isfinite(got).all())
    rms = ((got - want).pow(2).mean().sqrt() / want.abs().max()).item()
    errs = dict(batch=rel(got, want), rms=rms, row=rel_rows(got, want), row_valid=rel_rows(got, want, s.valid_aa_idx))
    print("forward[%s] %s L%d d%d %s: batch max %.3e  rms %.3e  worst row %.3e  worst row over valid ids %.3e"
          % (precision, arch, layers, d, shape, errs["batch"], rms, errs["row"], errs["row_valid"]))
    return errs


@pytest.mark.parametrize("arch,layers,d,H,F,shape", FORWARD_CASES)
def test_forward_logits_vs_oracle(arch, layers, d, H, F, shape):
    """Default ("fast") numerics: one pass of fp16 operands on the tensor cores, fp32 everywhere else.
    Tolerances -- north_star's 1e-3, measured two ways: over the batch (max|d| / max|logit|) and per token row (SURVEY
    8c).  Models of up to 12 layers (MSA-1b included) are inside it.  The 33-layer models sit AT the single-pass limit
    (DESIGN.md section 3): only the documented single-pass bounds are asserted for them here, and the 1e-3 tolerance
    itself is asserted in split-operand mode (test_forward_logits_precise_mode)."""
    e = _forward_errors(arch, layers, d, H, F, shape)
    assert e["rms"] < 0.5 * LOGIT_TOL
    if layers < 30:
        assert e["batch"] < LOGIT_TOL and e["row"] < 2 * LOGIT_TOL
    else:   # 33 layers, single pass: AT the tolerance, not inside it (BASELINE.md section 2 says so)
        assert e["batch"] < BATCH_TOL_FAST_DEEP and e["row"] < ROW_TOL_FAST


@pytest.mark.parametrize("arch,layers,d,H,F,shape", [
    ("roberta_large", 33, 1280, 20, 5120, (2, 66)), ("esm2", 33, 1280, 20, 5120, (2, 40)),
    ("roberta_large", 33, 1280, 20, 5120, (2, 258)), ("esm2", 33, 1280, 20, 5120, (2, 514)),     # configs 2 and 4
    ("msa_transformer", 12, 768, 12, 3072, (1, 6, 40)), ("esm1", 6, 768, 12, 3072, (2, 66)),
    ("esm2", 6, 320, 20, 1280, (2, 27)), ("msa_transformer", 2, 128, 4, 256, (1, 3, 70)),
])
def test_forward_logits_precise_mode(arch, layers, d, H, F, shape):
    """Split-operand mode (`precision="split"`, pgibbs_set_precision level 2): every GEMM runs
    a_hi w_hi + a_hi w_lo + a_lo w_hi on fp16 hi / lo pairs, so only the attention-internal roundings (q, k, v, P in
    fp16) remain.  north_star's 1e-3 holds per token row as well as over the batch for the 33-layer models, ESM-2 650M
    included -- the tolerance the single-pass mode cannot reach at that depth (CPU emulation: precision_study3.py)."""
    e = _forward_errors(arch, layers, d, H, F, shape, precision="split")
    assert e["batch"] < LOGIT_TOL and e["row"] < LOGIT_TOL and e["rms"] < 0.25 * LOGIT_TOL


@pytest.mark.parametrize("arch,d,H,shape", [("esm2", 128, 2, (72, 258)),        # 18 576 rows, d = 128: four rows per warp
                                            ("esm2", 640, 10, (20, 258)),       # 5 160 rows, d = 640: two rows per warp
                                            ("msa_transformer", 128, 2, (2, 60, 129))])
def test_lm_head_rows_per_warp_invariance(arch, d, H, shape, monkeypatch):
    """The LM-head kernel lets a warp carry 1, 2 or 4 rows through one pass over the projection table, chosen from the
    number of sampled rows; the per-row arithmetic must not depend on that choice (a chain's logits are the same in
    every batch), and every variant must agree with the oracle."""
    from oracle.fair_esm import OracleModel
    from protein_gibbs_sampler_b200.config import tiny_config
    cfg = tiny_config(arch, 2, d, H, 2 * d)
    tok = _tokens(cfg, shape, 9)
    got = {}
    for cap in ("1", "2", "4"):
        monkeypatch.setenv("PGIBBS_HEAD_ROWS", cap)
        s, sd = make(cfg, 4)
        got[cap] = s.model.model(tok)["logits"]
    assert torch.equal(got["1"], got["2"]) and torch.equal(got["1"], got["4"])
    want = OracleModel(cfg, sd).model(tok)["logits"]
    assert rel(got["4"], want) < LOGIT_TOL


@pytest.mark.parametrize("arch,shape", [("esm2", (2, 514)), ("roberta_large", (2, 258))])   # configs 4 and 2
def test_forward_logits_split_weights_mode_config_shapes(arch, shape):
    """The middle level (`precision="split_weights"`: weights as fp16 hi + lo pairs, activations single fp16; two passes
    per GEMM): the 650M models at the token shapes of configs 4 and 2 are inside north_star's 1e-3 on the batch metric
    (the per-row metric needs the activations' lo halves as well, i.e. "split")."""
    e = _forward_errors(arch, 33, 1280, 20, 5120, shape, precision="split_weights")
    assert e["batch"] < LOGIT_TOL and e["row"] < ROW_TOL_FAST and e["rms"] < 0.4 * LOGIT_TOL


def test_precision_levels_order_and_agree():
    """fast -> split_weights -> split: the error against the fp32 oracle shrinks level by level on the same inputs,
    and every level samples from the same chain state machinery (one generate runs in each mode)."""
    cfg_args = ("roberta_large", 12, 512, 8, 2048, (2, 130))
    errs = {pr: _forward_errors(*cfg_args, precision=pr) for pr in ("fast", "split_weights", "split")}
    assert errs["split"]["rms"] < 0.6 * errs["split_weights"]["rms"] and errs["split_weights"]["rms"] < 0.85 * errs["fast"]["rms"]
    from protein_gibbs_sampler_b200.config import tiny_config
    cfg = tiny_config("esm2", 2, 128, 2, 256)
    outs = {}
    for pr in ("fast", "split"):
        s, _ = make(cfg, 7, precision=pr)
        random.seed(11); torch.manual_seed(11)
        outs[pr] = s.generate(3, "MKTAYIAKQRQISFVKSHFSRQ", batch_size=3, num_iters=17, top_k=3, burnin=2, num_positions=5,
                              show_progress_bar=False)
        assert len(outs[pr]) == 3 and all(len(x) == 22 for x in outs[pr])
    same = sum(a == b for x, y in zip(outs["fast"], outs["split"]) for a, b in zip(x, y))
    assert same >= 0.8 * 66   # same noise, logits 1e-3 apart: nearly every draw agrees


def test_forward_outlier_weights_stay_finite_and_in_tolerance():
    """Real ESM checkpoints carry a few residual-stream channels and FFN units far above the rest; synthetic
    N(0, 0.02) weights never exercise the dynamic range of `h` / `qkv` / `ffn`.  Here four residual channels are fed
    by FC2 / out-projection rows scaled x60 and eight FC1 units are scaled x100 in four of six layers: the residual
    stream reaches the hundreds and the FFN activations the tens (fp16 range: 65504).
      * Logits stay finite in every mode.
      * Single-pass fp16 operands lose accuracy here, as any 16-bit inference does: after LayerNorm the outlier
        channels are ~300x the rest, and their 2^-12 relative rounding is an absolute error comparable to the SIGNAL
        of the small channels.  Documented bound: 1e-2 of the largest logit (measured 3.6e-3).
      * Split-operand mode carries every GEMM operand as hi + lo (22 bits); what remains is the fp16 rounding of
        q / k / v and P inside the attention, amplified by the same outliers: 2e-3 documented (measured 1.1e-3).
      * Beyond the fp16 range the engine saturates (cvt.satfinite) instead of producing inf / NaN."""
    from oracle.fair_esm import OracleModel
    from protein_gibbs_sampler_b200 import models
    from protein_gibbs_sampler_b200.config import tiny_config
    from protein_gibbs_sampler_b200.esm_sampler import ESM_sampler
    from protein_gibbs_sampler_b200.weights import synthetic_state_dict
    for arch in ("roberta_large", "esm2"):
        cfg = tiny_config(arch, 6, 320, 5, 1280)
        sd = synthetic_state_dict(cfg, 9)
        chans, units = [5, 77, 130, 301], list(range(40, 48))
        for li in (1, 3, 4, 5):
            sd["layers.%d.fc2.weight" % li][chans] *= 60.0
            sd["layers.%d.self_attn.out_proj.weight" % li][chans] *= 60.0
            sd["layers.%d.fc1.weight" % li][units] *= 100.0
            sd["layers.%d.fc1.bias" % li][units] *= 100.0
        tok = _tokens(cfg, (2, 70), 5)
        want = OracleModel(cfg, sd).model(tok)["logits"]
        for precision, tol in (("fast", 1e-2), ("split", 2 * LOGIT_TOL)):
            s = ESM_sampler(models.CustomModel(cfg, state_dict=sd, precision=precision), device="cuda:0")
            got = s.model.model(tok)["logits"]
            eng = s.model.model.engine
            pitch = 2 if precision == "split" else 1      # split mode: ffn rows are [hi | lo]
            x = eng.debug_read("x", tok.numel() * 320)
            ffn = eng.debug_read("ffn", tok.numel() * 1280 * pitch).view(tok.numel(), pitch, 1280)[:, 0]
            print("outliers %s [%s]: max|x| %.0f max|ffn| %.0f  batch %.3e row %.3e" % (
                arch, precision, x.abs().max(), ffn.abs().max(), rel(got, want), rel_rows(got, want)))
            assert x.abs().max() > 100 and ffn.abs().max() > 10        # the outliers are really there
            assert bool(torch.isfinite(got).all())
            assert rel(got, want) < tol, (arch, precision, rel(got, want))
        # beyond the fp16 range: saturate, never inf / NaN
        sd2 = {k: v.clone() for k, v in sd.items()}
        sd2["layers.2.fc1.weight"][units] *= 1.0e5
        s2 = ESM_sampler(models.CustomModel(cfg, state_dict=sd2), device="cuda:0")
        assert bool(torch.isfinite(s2.model.model(tok)["logits"]).all())


# ------------------------------------------------------------------------------------------ end to end
def test_generate_reproduces_reference_golden(golden):
    """Replay mode, same Python/torch seeds as the reference run: positions are bit-identical, and each
    iteration started from the reference's state yields the reference's next state (a logit difference below
    tolerance can only flip a draw that sits inside the error band, so allow <= 1 % of residues to differ)."""
    import numpy as np
    from oracle.fair_esm import OracleModel
    for c in golden["cases"]:
        kw = c["kwargs"]
        s, sd = make(c["cfg"], c["weights_seed"])
        random.seed(c["rng_seed"])
        torch.manual_seed(c["rng_seed"])
        out = s.generate(**kw)
        assert len(out) == len(c["output"]) and all(len(a) == len(b) for a, b in zip(out, c["output"]))
        same = sum(x == y for a, b in zip(out, c["output"]) for x, y in zip(a, b))
        total = sum(len(a) for a in c["output"])
        assert same >= 0.9 * total, (out, c["output"])
    exact = 0
    for c in golden["cases"]:
        random.seed(c["rng_seed"])
        torch.manual_seed(c["rng_seed"])
        s, sd = make(c["cfg"], c["weights_seed"])
        exact += s.generate(**c["kwargs"]) == c["output"]
    assert exact >= len(golden["cases"]) - 1


def _check_trace(sampler, oracle, case, n_rows_per_unit):
    """One golden case, iteration by iteration FROM THE REFERENCE'S OWN STATE: the token tensor the reference fed to
    forward number f (`states[f]`, already masked) is loaded into the engine, the reference's target positions of that
    iteration and the Exp(1) variates it consumed there are replayed, and ONE iteration runs on the device.  Then
      (1) every sampled residue equals the oracle's sampler tail applied to the ENGINE's logits of that state -- the
          device tail is bit-exact given its logits -- and the untouched positions are unchanged;
      (2) the engine's logits are within tolerance of the oracle's;
      (3) the oracle's tail on the ORACLE's logits reproduces the reference's next state (the trace is understood);
    so a residue that differs from the reference's can only come from a logit difference inside the tolerance band.
    Returns (residues sampled, residues that differ from the reference)."""
    from oracle.sampler_tail import generate_step_with_noise
    from protein_gibbs_sampler_b200.esm_sampler import draw_replay_noise, _effective_k
    kw = case["kwargs"]
    eng = sampler.model.model.require_engine()
    valid = sampler.valid_aa_idx
    num_iters, top_k, burnin = kw["num_iters"], kw.get("top_k", 0), kw.get("burnin", float("inf"))
    temperature = kw.get("temperature")
    states, targets = case["states"], case["targets"]
    if not targets:   # mask=False runs never call the (traced) masking hook: num_positions=0 -> every candidate position
        assert not kw.get("mask", True) and not kw.get("num_positions") and not kw.get("leader_length")
        T = torch.tensor(states[0]).shape[-1]
        L = T - 2 if n_rows_per_unit == 1 else T - 1          # <cls> seq <eos>  /  MSA rows: <cls> seq
        every = list(range(1, L + 1))
        targets = [[([every] * n_rows_per_unit if n_rows_per_unit > 1 else every) for _ in st] for st in states]
    assert len(states) == len(targets) and len(states) % num_iters == 0
    random.seed(case["rng_seed"]); torch.manual_seed(case["rng_seed"])
    sampled = differ = 0
    for f, (state, tg) in enumerate(zip(states, targets)):
        it = f % num_iters
        tok = torch.tensor(state)
        flat_t = [t for unit in tg for t in (unit if n_rows_per_unit > 1 else [unit])]   # per chain (MSA: per row)
        n_chains, P = len(flat_t), len(flat_t[0])
        assert n_chains == tok.numel() // tok.shape[-1]
        if it == 0:   # the reference draws from torch's generator residue by residue; one pre-draw per outer batch
            noise, stride = draw_replay_noise(num_iters, n_chains * P, len(valid), top_k, burnin)
        pos = torch.tensor(flat_t, dtype=torch.int32).reshape(1, n_chains, P)
        eng.set_tokens(tok)
        eng.set_schedule(pos.numpy(), 1, P, n_chains * P, P, False)
        eng.set_noise(noise[it:it + 1].contiguous(), stride)
        eng.set_chain_offset(0)
        eng.run(0, 1, float("inf") if it < burnin else 0, top_k, temperature, False, valid)
        got = eng.get_tokens().reshape(n_chains, -1)
        logits_e = eng.forward_logits(tok).reshape(n_chains, tok.shape[-1], -1)
        logits_o = oracle.model(tok)["logits"].reshape(n_chains, tok.shape[-1], -1)
        assert rel(logits_e, logits_o) < LOGIT_TOL
        before = tok.reshape(n_chains, -1)
        touched = torch.zeros_like(before, dtype=torch.bool)
        # the reference's next state: states[f+1] outside its own freshly masked targets, or the final output
        last = it == num_iters - 1
        nxt = None if last else torch.tensor(states[f + 1]).reshape(n_chains, -1)
        nxt_masked = set() if last or not kw.get("mask", True) else \
            {(c, p) for c, ps in enumerate(t for unit in targets[f + 1] for t in (unit if n_rows_per_unit > 1 else [unit]))
             for p in ps}
        k = _effective_k(top_k, len(valid), it < burnin)
        for c in range(n_chains):
            for j, p in enumerate(flat_t[c]):
                touched[c, p] = True
                nz = noise[it, c * P + j]
                mine = generate_step_with_noise(logits_e[c, p], nz, valid, k, temperature)
                assert int(got[c, p]) == mine, (f, c, p)                                   # (1)
                ref_tok = generate_step_with_noise(logits_o[c, p], nz, valid, k, temperature)
                if nxt is not None and (c, p) not in nxt_masked and flat_t[c].count(p) == 1:
                    assert int(nxt[c, p]) == ref_tok, (f, c, p)                            # (3)
                sampled += 1
                differ += mine != ref_tok
        assert torch.equal(got[~touched], before[~touched])                               # (1) untouched positions
    return sampled, differ


def test_generate_traces_iteration_by_iteration(golden):
    """The `states` / `targets` traces of the reference runs (tests/golden/make_golden.py), replayed one iteration at a
    time from the reference's own state, single-sequence and MSA samplers."""
    from oracle.fair_esm import OracleModel
    total = wrong = 0
    for c in golden["cases"] + golden["msa_cases"]:
        s, sd = make(c["cfg"], c["weights_seed"])
        msa = c["cfg"]["arch"] == "msa_transformer"
        n, d = _check_trace(s, OracleModel(c["cfg"], sd), c, len(c["kwargs"]["seed_msa"]) if msa else 1)
        total += n
        wrong += d
    print("trace replay: %d residues sampled, %d differ from the reference (each inside the logit error band)" % (total, wrong))
    assert total > 400 and wrong <= 0.02 * total


def test_msa_generate_reproduces_reference_golden(golden):
    for c in golden["msa_cases"]:
        s, sd = make(c["cfg"], c["weights_seed"])
        random.seed(c["rng_seed"])
        torch.manual_seed(c["rng_seed"])
        out = s.generate(**c["kwargs"])
        assert len(out) == len(c["output"])
        same = sum(x == y for a, b in zip(out, c["output"]) for x, y in zip(a, b))
        assert same >= 0.9 * sum(len(a) for a in c["output"]), (out, c["output"])
    for c in golden["single_cases"]:
        s, sd = make(c["cfg"], c["weights_seed"])
        random.seed(c["rng_seed"])
        torch.manual_seed(c["rng_seed"])
        out = s.generate_single(**c["kwargs"])
        same = sum(x == y for x, y in zip(out, c["output"]))
        assert len(out) == len(c["output"]) and same >= 0.8 * len(out), (out, c["output"])


@pytest.mark.parametrize("batch_size,num_positions,mask,leader_length,in_order", [
    (3, 1, True, 1, True), (3, 1, False, 1, True), (3, 1, True, 1, False), (3, 1, False, 1, False),
    (3, 1, True, -1, False), (10, 3, False, 1, False)])
def test_generate_invariants(batch_size, num_positions, mask, leader_length, in_order):
    """The reference's integration tests (test_esm_sampler.py:90-124): count, length, alphabet."""
    from protein_gibbs_sampler_b200.config import tiny_config
    s, _ = make(tiny_config("esm2", 2, 128, 2, 256), 0, rng="device")
    out = s.generate(4, "AAAAAAAAAA", batch_size=batch_size, max_len=10, num_iters=2, num_positions=num_positions,
                     mask=mask, leader_length=leader_length, in_order=in_order, show_progress_bar=False)
    assert len(out) == 4 and all(len(x) == 10 and set(x) <= set("ACDEFGHIKLMNPQRSTVWY") for x in out)
    out = s.generate(4, "", batch_size=4, max_len=10, show_progress_bar=False)
    assert len(out) == 4 and all(len(x) == 10 for x in out)
    out = s.generate(4, "", batch_size=10, max_len=10, show_progress_bar=False)
    assert len(out) == 4


def test_in_order_single_iteration_keeps_untouched_columns():
    """test_esm_msa_sampler.py:113-122: one in-order iteration only rewrites the scheduled columns."""
    from protein_gibbs_sampler_b200.config import tiny_config
    s, _ = make(tiny_config("msa_transformer", 2, 128, 2, 256), 0, rng="device")
    msa = ["MKTAYIAKQR", "MKSAYLAKQR", "MRTAYIAKQQ"]
    out = s.generate(3, msa, batch_size=1, num_iters=1, in_order=True, num_positions=2, leader_length=3,
                     show_progress_bar=False)
    for a, b in zip(out, msa):
        assert len(a) == len(b)
        changed = [i for i, (x, y) in enumerate(zip(a, b)) if x != y]
        assert set(changed) <= {6, 7}                      # cursor quirk: starts at indexes[leader % len]
    single = s.generate_single(["AAAAAA", "AAAAAA", "GGGGGG"], steps=2, passes=2, burn_in=1)
    assert len(single) == 6 and set(single) <= set("-ACDEFGHIKLMNPQRSTVWY")


def test_full_size_properties_config2():
    """BASELINE config 2 geometry (ESM-1b 650M, 64 chains x L=256), where the CPU oracle is too slow:
    determinism, only scheduled positions change, chain independence (a sub-batch reproduces its chains)."""
    from protein_gibbs_sampler_b200 import models
    from protein_gibbs_sampler_b200.esm_sampler import ESM_sampler
    rng = random.Random(1)
    seeds = ["".join(rng.choice("ACDEFGHIKLMNPQRSTVWY") for _ in range(256)) for _ in range(64)]
    m = models.ESM1b(seed=0)
    s = ESM_sampler(m, device="cuda:0", rng="device")   # Philox keyed by (iteration, chain, slot)
    toks = m.batch_converter([(str(i), x) for i, x in enumerate(seeds)])[2]
    eng = m.model.engine
    positions = sorted(random.Random(2).sample(range(1, 257), 40))

    def run(tokens, n_iters=2):
        torch.manual_seed(7)
        plan, _ = s.plan_positions(tokens.shape[0], positions, -1, 0, False, n_iters)
        return s.run_plan(tokens, plan, top_k=3, temperature=None, burnin=1, mask=True)[:, 0]

    a = run(toks)
    b = run(toks)
    assert torch.equal(a, b)                                            # deterministic under a fixed seed
    untouched = [i for i in range(258) if i not in positions]
    assert torch.equal(a[:, untouched], toks[:, untouched])              # masking/indexing exact
    assert bool(((a[:, positions] >= 4) & (a[:, positions] <= 23)).all())   # only the 20 amino acids are written
    assert not torch.equal(a[:, positions], toks[:, positions])
    sub = run(toks[:16])
    assert torch.equal(sub, a[:16])                                      # chains are independent Markov chains
    logits = eng.forward_logits(toks[:2])
    assert logits.shape == (2, 258, 33) and bool(torch.isfinite(logits).all())


def test_generate_single_batch_equals_sequential_calls():
    """n generate_single chains as one device batch give what n sequential calls give (replay mode:
    shuffles and Exp(1) variates are pre-drawn in the sequential calls' order)."""
    from protein_gibbs_sampler_b200.config import tiny_config
    cfg = tiny_config("msa_transformer", 2, 128, 2, 256)
    s, _ = make(cfg, 3)
    msa = ["MKTAYIAK-RQ", "MKSAY-AKQRQ", "MRTAYIAKQ-Q", "M-TAYLAKQRQ"]
    for kw in (dict(steps=3, passes=2, burn_in=1, target_index=0, k=1),
               dict(steps=2, passes=3, burn_in=1, target_index=-1, k=2, exclude_positions=[0, 4]),
               dict(steps=11, passes=1, burn_in=0, target_index=2, k=3)):
        random.seed(5)
        torch.manual_seed(5)
        want = [s.generate_single(msa, **kw) for _ in range(3)]
        random.seed(5)
        torch.manual_seed(5)
        got = s.generate_single_batch(msa, 3, **kw)
        assert got == want, (kw, got, want)
        assert len(set(want)) > 1 or kw["k"] == 1   # the chains really are different draws


def test_esm1_generate_matches_oracle_loop():
    """ESM-1 family end to end (the reference's esm6 / esm12 / esm34): replay mode against the oracle port of the
    reference loop on the ESM-1 oracle model, same seeds -> same position schedule; residues may differ only where
    a draw sits inside the logit error band."""
    from oracle.fair_esm import OracleModel
    from oracle.gibbs_loop import esm_generate
    from protein_gibbs_sampler_b200.config import tiny_config
    cfg = tiny_config("esm1", 2, 128, 2, 256)
    s, sd = make(cfg, 11)
    kw = dict(seed_seq="MKTAYIAKQRQISFVKSHFSRQ", batch_size=3, num_iters=3, top_k=3, burnin=1, num_positions=5)
    random.seed(3); torch.manual_seed(3)
    want = esm_generate(OracleModel(cfg, sd), 3, **kw)
    random.seed(3); torch.manual_seed(3)
    got = s.generate(3, show_progress_bar=False, **kw)
    assert len(got) == 3 and all(len(x) == 22 and set(x) <= set("ACDEFGHIKLMNPQRSTVWY") for x in got)
    diff = sum(a != b for x, y in zip(got, want) for a, b in zip(x, y))
    assert diff <= 2, (got, want)


# ------------------------------------------------------------------------------------------------ scoring
@pytest.mark.parametrize("arch,H,length,rows,mask_distance,batch_size", [
    ("roberta_large", 2, 23, 1, float("inf"), None), ("roberta_large", 2, 23, 1, 5, 2), ("esm2", 2, 131, 1, 7, 3),
    ("esm1", 2, 30, 1, 4, None), ("roberta_large", 2, 20, 1, None, None),
    ("msa_transformer", 2, 19, 4, 6, 4), ("msa_transformer", 2, 19, 3, float("inf"), 5), ("msa_transformer", 2, 19, 3, None, 1),
])
def test_device_scoring_equals_logits_path(arch, H, length, rows, mask_distance, batch_size):
    """Engine.score (on-device strided <mask>, LM head on the masked rows, fused log_softmax + gather) against the
    same quantity computed on the host from forward_logits of host-masked copies, as the reference's
    log_likelihood_batch does it (esm_sampler.py:316-352, esm_msa_sampler.py:371-424).  Same forward, so only the
    fp32 log_softmax differs: 2e-5 absolute.  mask_distance None = --masking_off."""
    from protein_gibbs_sampler_b200.config import tiny_config
    from protein_gibbs_sampler_b200.esm_sampler import ESM_ALLOWED_AMINO_ACIDS
    cfg = tiny_config(arch, 2, 128, H, 256)
    s, _ = make(cfg, 11)
    rnd = random.Random(length * 7 + rows)
    seqs = ["".join(rnd.choice(ESM_ALLOWED_AMINO_ACIDS) for _ in range(length)) for _ in range(rows)]
    masking = mask_distance is not None
    md = mask_distance if masking else float("inf")
    alphabet = s.model.alphabet
    start = 1 if alphabet.prepend_bos else 0
    engine = s.model.model.require_engine()
    if arch == "msa_transformer":
        seqs[1] = "-" + seqs[1][1:5] + "--" + seqs[1][7:]          # gaps in the target row are skipped
        target = 1
        mean, each = s.log_likelihood(seqs, target_index=target, with_masking=masking, mask_distance=md)
        toks = s.model.batch_converter([[(str(i), q) for i, q in enumerate(seqs)]])[2]
        true = toks[0, target]
        scored = [i for i in range(length) if seqs[target][i] != "-"]
    else:
        mean, each = next(s.log_likelihood_batch(seqs[:1], with_masking=masking, mask_distance=md, batch_size=batch_size))
        toks = s.model.batch_converter([("0", seqs[0])])[2]
        true = toks[0]
        scored = list(range(length))
    n = int(min(md, length)) if masking else 1
    want = {}
    for i in range(n):
        t = toks.clone()
        row = t[0, target] if arch == "msa_transformer" else t[0]
        if masking:
            row[start + i:start + length:n] = alphabet.mask_idx
        lp = torch.log_softmax(engine.forward_logits(t), dim=-1)[0]
        lp = lp[target] if arch == "msa_transformer" else lp
        for pos in range(i, length, n):
            want[pos] = lp[start + pos, true[start + pos]].item()
    order = [pos for i in range(n) for pos in range(i, length, n) if pos in scored]
    assert len(each) == len(order)
    assert each == pytest.approx([want[pos] for pos in order], abs=2e-5)
    assert mean == pytest.approx(sum(want[pos] for pos in order) / len(order), abs=2e-5)


# --------------------------------------------------------------------------------------- CUDA-graph replay
@pytest.mark.parametrize("arch,rng", [("esm2", "replay"), ("roberta_large", "device"), ("esm1", "replay"),
                                      ("msa_transformer", "replay"), ("msa_transformer", "device")])
def test_graph_replay_equals_kernel_by_kernel(arch, rng, monkeypatch):
    """Runs of >= 16 iterations launch the first iteration kernel by kernel and replay a captured CUDA graph for the
    rest, with the iteration index (schedule slice, noise slice, burn-in switch of top_k, RNG counter) read on the
    device.  Same seeds -> the same sequences as PGIBBS_GRAPH=0, bit for bit."""
    from protein_gibbs_sampler_b200.config import tiny_config
    cfg = tiny_config(arch, 2, 128, 2, 256)
    rnd = random.Random(5)
    seq = "".join(rnd.choice("ACDEFGHIKLMNPQRSTVWY") for _ in range(30))
    kw = dict(num_iters=19, burnin=6, top_k=3, num_positions=7, show_progress_bar=False)
    outs = []
    for graph in ("0", "1"):
        monkeypatch.setenv("PGIBBS_GRAPH", graph)
        s, _ = make(cfg, 7, rng=rng)
        random.seed(11); torch.manual_seed(11)
        if arch == "msa_transformer":
            msa = [seq, seq[:10] + "A" + seq[11:], seq[:20] + "-" + seq[21:]]
            outs.append(s.generate(4, msa, batch_size=2, **kw))
            random.seed(12); torch.manual_seed(12)
            outs[-1] = outs[-1] + [s.generate_single(msa, steps=17, passes=2, burn_in=1)]
        else:
            outs.append(s.generate(5, seq, batch_size=3, in_order=(arch == "esm1"), **kw))
    assert outs[0] == outs[1]
    assert len(set(outs[1])) > 1   # the chains did move


@pytest.mark.parametrize("arch,d,H,shape", [("roberta_large", 512, 8, (40, 258)),       # 82 tiles on 74 CTA pairs: 74 + 8
                                            ("esm2", 512, 8, (41, 257)),
                                            ("msa_transformer", 768, 12, (4, 20, 129))])  # 123 tiles: 74 + 49
def test_tail_overlap_bit_identical(arch, d, H, shape, monkeypatch):
    """PGIBBS_TAIL_OVERLAP=1 cuts every residual GEMM at its last full wave and runs the LayerNorm of the finished row
    blocks on a second stream next to the partly-filled last wave (fork / join; parallel graph branches under capture).
    Same kernels on the same rows: logits and a 17-iteration graph-replayed generate must equal the plain order bit for bit."""
    from protein_gibbs_sampler_b200.config import tiny_config
    cfg = tiny_config(arch, 3, d, H, 4 * d)
    tok = _tokens(cfg, shape, 21)
    logits, seqs = [], []
    for mode in ("0", "1"):
        monkeypatch.setenv("PGIBBS_TAIL_OVERLAP", mode)
        s, _ = make(cfg, 9)
        logits.append(s.model.model(tok)["logits"])
        if arch != "msa_transformer":
            rnd = random.Random(3)
            seq = "".join(rnd.choice("ACDEFGHIKLMNPQRSTVWY") for _ in range(shape[1] - 2))
            random.seed(4); torch.manual_seed(4)
            seqs.append(s.generate(shape[0], seq, batch_size=shape[0], num_iters=17, burnin=5, top_k=3, num_positions=9,
                                   show_progress_bar=False))
    assert bool(torch.isfinite(logits[0]).all()) and torch.equal(logits[0], logits[1])
    if seqs:
        assert seqs[0] == seqs[1] and len(set(seqs[1])) > 1


# ----------------------------------------------------------------- full-size properties of the other BASELINE configs
def _full_size_properties(sampler, toks, plan, rows_per_unit, valid_hi, units_sub, **run_kw):
    """Determinism, only scheduled positions change, only valid residues are written, and a sub-batch of chains /
    MSAs with its slice of the schedule reproduces itself (independent Markov chains: what sharding across GPUs
    relies on)."""
    def run(tokens, pl):
        torch.manual_seed(7)    # the device RNG seed is drawn from torch's generator
        return sampler.run_plan(tokens, pl, mask=True, **run_kw)

    n_chains = toks.shape[0] * rows_per_unit
    a = run(toks, plan)
    assert torch.equal(a, run(toks, plan))
    flat_in, flat_out = toks.reshape(n_chains, -1), a.reshape(n_chains, -1)
    touched = torch.zeros_like(flat_in, dtype=torch.bool)
    for it in range(plan.n_iters):
        for c in range(n_chains):
            touched[c, plan.targets(it, c)] = True
    assert torch.equal(flat_out[~touched], flat_in[~touched])
    assert bool(((flat_out[touched] >= 4) & (flat_out[touched] <= valid_hi)).all())
    assert not torch.equal(flat_out[touched], flat_in[touched])
    sub = run(toks[:units_sub], plan.slice_chains(0, units_sub * rows_per_unit, n_chains))
    assert torch.equal(sub, a[:units_sub])
    return a


def test_full_size_properties_config3_msa():
    """BASELINE config 3 geometry: MSA-1b, 16 MSAs x 32 rows x L=128, 10 % of the positions per row and iteration."""
    from protein_gibbs_sampler_b200 import models
    from protein_gibbs_sampler_b200.esm_msa_sampler import ESM_MSA_sampler
    rng = random.Random(1)
    rows = ["".join(rng.choice("ACDEFGHIKLMNPQRSTVWY") for _ in range(128)) for _ in range(32)]
    m = models.ESM_MSA1(seed=0)
    s = ESM_MSA_sampler(m, device="cuda:0", rng="device")
    toks = m.batch_converter([[(str(i), x) for i, x in enumerate(rows)]] * 16)[2]
    idx, _ = s.calculate_indexes(None, 0, 128, False)
    random.seed(3)
    plan, _ = s.plan_positions(16, 32, idx, -1, 12, False, 2)
    out = _full_size_properties(s, toks, plan, 32, 30, 4, top_k=0, temperature=None, burnin=float("inf"))
    assert out.shape == (16, 32, 129)


@pytest.mark.parametrize("name,B,L,P,top_k", [("esm2_t33_650M", 64, 512, 0, 5), ("esm1b", 16, 1022, 51, 0)])
def test_full_size_properties_config4_and_5_shards(name, B, L, P, top_k):
    """Per-GPU shards of BASELINE configs 4 (ESM-2 650M, 64 of 512 chains x L=512, top_k 5) and 5 (ESM-1b, 16 of 128
    chains x L=1022 = the longest sequence the learned positions allow, 5 % of the positions per iteration)."""
    from protein_gibbs_sampler_b200 import models
    from protein_gibbs_sampler_b200.esm_sampler import ESM_sampler
    rng = random.Random(1)
    seeds = ["".join(rng.choice("ACDEFGHIKLMNPQRSTVWY") for _ in range(L)) for _ in range(B)]
    m = (models.ESM2_t33_650M if name == "esm2_t33_650M" else models.ESM1b)(seed=0)
    s = ESM_sampler(m, device="cuda:0", rng="device")
    toks = m.batch_converter([(str(i), x) for i, x in enumerate(seeds)])[2]
    idx, last_i = s.calculate_indexes(None, 0, L, False)
    if not P:   # config 4 resamples every position; keep a fixed subset here so that some positions must stay put
        idx = sorted(random.Random(2).sample(list(idx), 64))
    random.seed(3)
    plan, _ = s.plan_positions(B, idx, last_i, P, False, 2)
    out = _full_size_properties(s, toks[:, None], plan, 1, 23, B // 4, top_k=top_k, temperature=None, burnin=1)
    assert out.shape == (B, 1, L + 2)


def test_two_engines_on_two_devices_in_one_process():
    """`pgibbs.h`: one engine per GPU, several per process.  Kernel attributes (the opt-in to > 48 KB of dynamic shared
    memory) are per device: an engine created on cuda:1 AFTER one on cuda:0 must configure its own device (a
    process-wide "configured" flag used to skip it -> invalid-argument launch failures).  Both engines give the same
    logits for the same weights and tokens, and each keeps working after the other has been used."""
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    from protein_gibbs_sampler_b200 import models
    from protein_gibbs_sampler_b200.config import tiny_config
    from protein_gibbs_sampler_b200.esm_sampler import ESM_sampler
    from protein_gibbs_sampler_b200.weights import synthetic_state_dict
    cfg = tiny_config("roberta_large", 2, 256, 4, 512)      # head_dim 64: tcgen05 GEMM + attention + LM-head kernels
    sd = synthetic_state_dict(cfg, 4)
    tok = _tokens(cfg, (2, 258), 5)
    s0 = ESM_sampler(models.CustomModel(cfg, state_dict=sd), device="cuda:0")
    a0 = s0.model.model(tok)["logits"]
    s1 = ESM_sampler(models.CustomModel(cfg, state_dict=sd), device="cuda:1")
    a1 = s1.model.model(tok)["logits"]
    assert torch.equal(a0, a1)
    assert torch.equal(s0.model.model(tok)["logits"], a0)
    random.seed(1); torch.manual_seed(1)
    out1 = s1.generate(2, "MKTAYIAKQRQISFVKSHFSRQ", batch_size=2, num_iters=2, top_k=3, show_progress_bar=False)
    random.seed(1); torch.manual_seed(1)
    out0 = s0.generate(2, "MKTAYIAKQRQISFVKSHFSRQ", batch_size=2, num_iters=2, top_k=3, show_progress_bar=False)
    assert out0 == out1
