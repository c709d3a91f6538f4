"""Stage-by-stage GPU bring-up check.  Each stage runs in its own subprocess (a device trap poisons the CUDA
context) with a timeout, so one bad kernel cannot hide the others.  Usage (on a GPU box):
    python tests/tools/gpu_check.py [stage ...]      # writes gpurun_out/gpu_check.log
"""
import os
import subprocess
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)


def rel_err(got, want):
    return ((got - want).abs().max() / want.abs().max().clamp_min(1e-30)).item()


def describe_mismatch(got, want, tol):
    import torch
    d = (got - want).abs()
    bad = d > tol * want.abs().max()
    n_bad = int(bad.sum())
    msg = "max_abs=%.4e rel=%.4e bad=%d/%d" % (d.max().item(), rel_err(got, want), n_bad, got.numel())
    if n_bad:
        rows = bad.any(1).nonzero().flatten()
        cols = bad.any(0).nonzero().flatten()
        msg += " bad_rows[%d]: %s.. bad_cols[%d]: %s.." % (len(rows), rows[:8].tolist(), len(cols), cols[:8].tolist())
        i = bad.nonzero()[0]
        msg += " first(%d,%d): got %.5f want %.5f" % (i[0], i[1], got[i[0], i[1]], want[i[0], i[1]])
        msg += " nan=%d" % int(torch.isnan(got).sum())
    return msg


def stage_gemm_pattern():
    import torch
    from protein_gibbs_sampler_b200.engine import op_gemm
    M, N, K = 128, 64, 64
    A = torch.zeros(M, K)
    A[torch.arange(M), torch.arange(M) % K] = 1.0
    for name, B in (("B=n", torch.arange(N).float()[:, None].expand(N, K).contiguous()),
                    ("B=k", torch.arange(K).float()[None, :].expand(N, K).contiguous())):
        C = op_gemm(A, B, epilogue=5, block_n=64)
        want = A @ B.t()
        print("pattern", name, describe_mismatch(C, want, 1e-3))
        if name == "B=k":
            print("  row0[:8]", C[0, :8].tolist(), "row5[:4]", C[5, :4].tolist(), "row70[:4]", C[70, :4].tolist())
        else:
            print("  row0[:8]", C[0, :8].tolist(), "row0[56:64]", C[0, 56:64].tolist())


def stage_gemm():
    import torch
    from protein_gibbs_sampler_b200.engine import op_gemm
    torch.manual_seed(0)
    ok = True
    cases = [(128, 64, 64, 64), (128, 256, 64, 256), (128, 128, 128, 128), (256, 192, 256, 192),
             (300, 320, 320, 64), (516, 960, 320, 192), (1000, 1280, 1280, 256), (129, 336, 128, 128)]
    for (M, N, K, bn) in cases:
        A = torch.randn(M, K) * 0.5
        B = torch.randn(N, K) * 0.5
        bias = torch.randn(N)
        want = A.half().float() @ B.half().float().t() + bias
        for epi, nm in ((5, "bias_f32"), (0, "bias_f16"), (1, "gelu_f16"), (4, "gelu_f32"), (2, "resid")):
            w = want
            C0 = None
            if epi in (1, 4):
                w = torch.nn.functional.gelu(want)
            if epi == 2:
                C0 = torch.randn(M, N)
                w = want + C0
            got = op_gemm(A, B, bias, C=C0, epilogue=epi, block_n=bn)
            tol = 2e-3 if epi in (0, 1) else 2e-5
            e = rel_err(got, w)
            good = e < tol
            ok &= good
            print("gemm M%d N%d K%d bn%d %-8s %s %s" % (M, N, K, bn, nm, "ok " if good else "FAIL",
                                                      describe_mismatch(got, w, tol)))
    print("GEMM_ALL_OK" if ok else "GEMM_HAS_FAILURES")


def stage_gemm_perf():
    import torch
    from protein_gibbs_sampler_b200.engine import op_gemm
    torch.manual_seed(0)
    for (M, N, K) in [(16512, 3840, 1280), (16512, 1280, 1280), (16512, 5120, 1280), (16512, 1280, 5120)]:
        A = torch.randn(M, K) * 0.1
        B = torch.randn(N, K) * 0.1
        for bn in (256, 192, 128):
            for epi in (5, 0, 1, 2):
                _, ms = op_gemm(A, B, torch.zeros(N), epilogue=epi, block_n=bn, reps=10)
                print("gemm_perf M%d N%d K%d bn%d epi%d: %.3f ms  %.1f TFLOP/s" % (M, N, K, bn, epi, ms,
                                                                                 2.0 * M * N * K / ms / 1e9))


def torch_attention(qkv, n_seq, T, H, Dh):
    import torch
    d = H * Dh
    x = qkv.half().float().view(n_seq, T, 3, H, Dh)
    q, k, v = x[:, :, 0].transpose(1, 2), x[:, :, 1].transpose(1, 2), x[:, :, 2].transpose(1, 2)
    p = torch.softmax(q @ k.transpose(-1, -2), dim=-1)
    return (p @ v).transpose(1, 2).reshape(n_seq * T, d)


def stage_attention():
    import torch
    from protein_gibbs_sampler_b200.engine import op_attention
    torch.manual_seed(0)
    ok = True
    for (n_seq, T, H, Dh) in [(2, 64, 2, 64), (2, 27, 20, 16), (3, 258, 4, 64), (2, 130, 3, 32), (1, 514, 2, 64)]:
        qkv = torch.randn(n_seq * T, 3 * H * Dh) * 0.7
        got = op_attention(qkv, n_seq, T, H, Dh)
        want = torch_attention(qkv, n_seq, T, H, Dh)
        e = rel_err(got, want)
        good = e < 3e-3
        ok &= good
        print("attn n%d T%d H%d Dh%d %s %s" % (n_seq, T, H, Dh, "ok " if good else "FAIL",
                                             describe_mismatch(got, want, 3e-3)))
    print("ATTN_ALL_OK" if ok else "ATTN_HAS_FAILURES")


def stage_sample():
    import torch
    from oracle.sampler_tail import generate_step_with_noise
    from protein_gibbs_sampler_b200.engine import op_sample
    torch.manual_seed(0)
    ok = True
    for valid, top_k, temp in [(list(range(4, 24)), 0, None), (list(range(4, 24)), 3, None),
                               (list(range(4, 24)) + [30], 5, 0.7), ([3, 5, 1], 2, None), (list(range(4, 24)), 1, 2.0)]:
        rows, V = 4000, 33
        logits = torch.randn(rows, V) * 2
        n = len(valid)
        noise = torch.empty(rows, n).exponential_(1)
        got = op_sample(logits, noise, valid, top_k=top_k, temperature=temp)
        want = torch.tensor([generate_step_with_noise(logits[i], noise[i], valid, top_k, temp) for i in range(rows)])
        mism = int((got != want).sum())
        ok &= mism == 0
        print("sample n_valid=%d top_k=%d temp=%s mismatches=%d/%d" % (n, top_k, temp, mism, rows))
    print("SAMPLE_ALL_OK" if ok else "SAMPLE_HAS_FAILURES")


def _forward_case(arch, layers, d, heads, ffn, B, T, R=1, seed=0, taps=True):
    import torch
    from oracle.fair_esm import OracleModel
    from protein_gibbs_sampler_b200 import models
    from protein_gibbs_sampler_b200.config import tiny_config
    from protein_gibbs_sampler_b200.weights import synthetic_state_dict
    cfg = tiny_config(arch, layers=layers, embed_dim=d, heads=heads, ffn_dim=ffn)
    sd = synthetic_state_dict(cfg, seed)
    g = torch.Generator().manual_seed(seed + 1)
    shape = (B, T) if arch != "msa_transformer" else (B, R, T)
    tok = torch.randint(4, 24, shape, generator=g)
    tok[..., 0] = 0
    if arch != "msa_transformer":
        tok[..., -1] = 2
    flat = tok.view(-1, T)
    flat[0, 3:9] = 32
    if flat.shape[0] > 1:
        flat[1, 1:T - 1] = 32
    taps_o = {}
    om = OracleModel(cfg, sd, hook=(lambda n, t: taps_o.__setitem__(n, t.detach().clone())))
    want = om.model(tok)["logits"]
    m = models.CustomModel(cfg, state_dict=sd)
    m.model.to("cuda:0")
    got = m.model(tok)["logits"]
    e = rel_err(got, want)
    print("forward %s L%d d%d H%d F%d B%d R%d T%d: rel=%.3e %s" % (arch, layers, d, heads, ffn, B, R, T, e,
                                                                  "ok" if e < 1e-3 else "FAIL"))
    if e >= 1e-3 and taps and arch != "msa_transformer":
        eng = m.model.engine
        M = tok.numel()
        for lim in range(0, layers + 1):
            eng.debug_layer_limit(lim)
            m.model(tok)
            x = eng.debug_read("x", M * d).view(M, d)
            ref = taps_o["embed" if lim == 0 else "layer%d" % (lim - 1)].reshape(M, d)
            print("   after %d layers: x rel=%.3e" % (lim, rel_err(x, ref)))
            if lim >= 1 and rel_err(x, ref) > 1e-2:
                qkv = eng.debug_read("qkv", M * 3 * d).view(M, 3 * d)
                ctx = eng.debug_read("ctx", M * d).view(M, d)
                print("   qkv absmax %.3f ctx absmax %.3f nan %d" % (qkv.abs().max(), ctx.abs().max(),
                                                                   int(torch.isnan(x).sum())))
                break
        eng.debug_layer_limit(-1)
    return e


def stage_forward():
    ok = True
    ok &= _forward_case("esm2", 2, 128, 2, 256, 2, 24) < 1e-3
    ok &= _forward_case("roberta_large", 2, 128, 2, 256, 2, 24) < 1e-3
    ok &= _forward_case("esm2", 6, 320, 20, 1280, 2, 27) < 1e-3
    ok &= _forward_case("roberta_large", 3, 256, 4, 512, 3, 130) < 1e-3
    ok &= _forward_case("esm2", 2, 640, 20, 2560, 2, 70) < 1e-3
    print("FORWARD_ALL_OK" if ok else "FORWARD_HAS_FAILURES")


def stage_forward_big():
    # full-depth ESM-1b geometry, small batch: the oracle finishes in seconds
    e = _forward_case("roberta_large", 33, 1280, 20, 5120, 2, 66, taps=False)
    e2 = _forward_case("esm2", 33, 1280, 20, 5120, 2, 66, taps=False)
    print("FORWARD_BIG_OK" if max(e, e2) < 1e-3 else "FORWARD_BIG_FAIL")


def stage_forward_msa():
    ok = _forward_case("msa_transformer", 2, 128, 2, 256, 2, 17, R=4) < 1e-3
    ok &= _forward_case("msa_transformer", 2, 768, 12, 3072, 1, 33, R=8) < 1e-3
    print("FORWARD_MSA_OK" if ok else "FORWARD_MSA_FAIL")


def stage_smoke():
    import __graft_entry__ as g
    g.smoke()
    print("SMOKE_OK")


STAGES = ["gemm_pattern", "gemm", "attention", "sample", "forward", "forward_big", "forward_msa", "smoke",
          "gemm_perf"]

if __name__ == "__main__":
    if len(sys.argv) >= 3 and sys.argv[1] == "--stage":
        globals()["stage_" + sys.argv[2]]()
        sys.exit(0)
    todo = sys.argv[1:] or STAGES
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    log = open(os.path.join(ROOT, "gpurun_out", "gpu_check.log"), "a")
    for st in todo:
        t0 = time.time()
        try:
            r = subprocess.run([sys.executable, os.path.abspath(__file__), "--stage", st], capture_output=True,
                               text=True, timeout=600, cwd=ROOT)
            out = r.stdout + ("\n[stderr]\n" + r.stderr[-3000:] if r.returncode else "")
            status = "rc=%d" % r.returncode
        except subprocess.TimeoutExpired as ex:
            out = (ex.stdout or b"").decode() if isinstance(ex.stdout, bytes) else (ex.stdout or "")
            status = "TIMEOUT"
        msg = "===== stage %s: %s (%.1fs)\n%s\n" % (st, status, time.time() - t0, out)
        print(msg)
        log.write(msg)
        log.flush()
