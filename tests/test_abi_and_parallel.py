"""CPU-only checks of the C-ABI library (loads, exports every declared symbol, fails loudly without a GPU)
and of the multi-rank plumbing (gloo, world_size 2)."""
import ctypes
import os
import re

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from protein_gibbs_sampler_b200 import _lib, parallel
from protein_gibbs_sampler_b200.config import get_config, tiny_config
from protein_gibbs_sampler_b200.weights import synthetic_state_dict

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    text = open(os.path.join(ROOT, "include", "pgibbs.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(pgibbs_[a-z_0-9]+)\s*\(", text)))


def test_library_exports_every_declared_symbol():
    lib = _lib.load()
    names = declared_symbols()
    assert len(names) >= 20
    for n in names:
        assert hasattr(lib, n), n
    assert set(names) == set(_lib.SIGNATURES), set(names) ^ set(_lib.SIGNATURES)
    assert b"sm_100a" in lib.pgibbs_version()


@pytest.mark.skipif(torch.cuda.is_available(), reason="CPU box behaviour")
def test_attention_kernel_keeps_its_loop_state_in_registers():
    """ptxas report of the last build (build.log, -Xptxas -v): the tcgen05 attention kernel must not use local memory.
    Round 2 measured what it costs when it does -- a never-executed debug timeline and a printf pushed the kernel to the
    168-register cap, the softmax loop re-loaded its loop-carried state from local memory in every hand-over: 12 % of
    the kernel (profiles/r02_attention_experiments.txt)."""
    log = os.path.join(ROOT, "protein_gibbs_sampler_b200", "build.log")
    if not os.path.exists(log):
        pytest.skip("no build.log (library built elsewhere)")
    lines = open(log).read().splitlines()
    hits = [i for i, l in enumerate(lines) if "Compiling entry function" in l and "attention_fa_kernel" in l]
    if not hits:
        pytest.skip("build.log has no ptxas report for the attention kernel")
    for i in hits:
        report = " ".join(lines[i:i + 4])
        m = re.search(r"(\d+) bytes stack frame, (\d+) bytes spill stores, (\d+) bytes spill loads", report)
        regs = re.search(r"Used (\d+) registers", report)
        assert m and regs, report
        assert (int(m.group(1)), int(m.group(2)), int(m.group(3))) == (0, 0, 0), report
        assert int(regs.group(1)) <= 160, report   # 384 threads: the cap is 168


@pytest.mark.skipif(torch.cuda.is_available(), reason="CPU box behaviour")
def test_gemm_accumulator_handback_carries_no_gpu_fence():
    """SASS of the built library (cuobjdump, no GPU): the CTA-pair GEMMs hand an accumulator stage back to the leader CTA
    with ONE SYNCS.ARRIVE per tile.  `mbarrier.arrive.release.cluster` put MEMBAR.ALL.GPU + ERRBAR + CGAERRBAR in front of it --
    lane 0 drained its outstanding TMA stores before every hand-back, 17-29 % of an epilogue warp's stall samples
    (profiles/r02zzz6_arrive_semantics_abc.txt).  Only the residual kernels may still contain a GPU
    fence (the opt-in K-split's flag protocol); the plain-store kernels also carry the 32-byte direct stores."""
    import shutil
    import subprocess
    if not shutil.which("cuobjdump"):
        pytest.skip("cuobjdump not on PATH")
    sass = subprocess.run(["cuobjdump", "-sass", _lib.LIB_PATH], capture_output=True, text=True).stdout
    per_kernel, name = {}, None
    for line in sass.splitlines():
        m = re.search(r"Function : (\S+)", line)
        if m:
            name = m.group(1)
            per_kernel[name] = []
        elif name:
            m = re.match(r"\s+/\*[0-9a-f]+\*/\s+(?:@!?U?P\d\s+)?([A-Z0-9_.]+)", line)
            if m:
                per_kernel[name].append(m.group(1))
    gemms = {k: v for k, v in per_kernel.items() if "gemm_tcgen05_kernel" in k}
    assert len(gemms) >= 12
    for k, ops in gemms.items():
        epi, cg = (int(x) for x in re.search(r"ILi\d+ELi(\d)ELi([12])E", k).groups())
        if epi != 2:                                   # everything but EPI_RESID_F32
            # what is left are the two barrier.cluster.arrive.release of a CTA pair (set-up and tear-down)
            fences = 2 if cg == 2 else 0
            assert ops.count("ERRBAR") == fences and ops.count("MEMBAR.ALL.GPU") == fences and ops.count("CGAERRBAR") == fences, k
            assert any(o.startswith("STG.E") and o.endswith(".256") for o in ops), k      # st.global.v8
        assert any(o.startswith("UTCHMMA") for o in ops), k


def test_tail_plan_matches_brute_force():
    """Where the engine cuts a residual GEMM for the LayerNorm overlap (engine.cu: tail_plan, through the host-only
    pgibbs_debug_tail_plan): the rows handed to the early LayerNorm must be exactly the row blocks whose every column tile
    lies in the full waves -- in both walking directions -- and the last wave must not touch them."""
    lib = _lib.load()
    out = (ctypes.c_int32 * 4)()
    cases = [(16512, 1280, 256, 2), (16512, 1280, 256, 1), (32896, 1280, 256, 2), (66048, 768, 256, 2), (10320, 512, 256, 2),
             (10537, 512, 128, 2), (10320, 768, 192, 2), (5000, 320, 64, 1), (258, 1280, 256, 2), (18944, 1280, 256, 2),
             (16384, 1280, 128, 1), (40000, 2560, 256, 2)]
    seen_overlap = 0
    for M, N, bn, cg in cases:
        for sms in (148, 132):
            for rev in (0, 1):
                _lib.check(lib.pgibbs_debug_tail_plan(M, N, bn, cg, sms, rev, out))
                full, rem, lo, hi = list(out)
                rpb = 128 * cg
                m_tiles, n_tiles = -(-M // rpb), -(-N // bn)
                tiles, groups = m_tiles * n_tiles, min(sms // cg, m_tiles * n_tiles)
                assert full + rem == tiles and full % groups == 0 and 0 <= rem < groups
                blk = lambda pos: ((tiles - 1 - pos) if rev else pos) // n_tiles
                count = {}
                for pos in range(full):
                    count[blk(pos)] = count.get(blk(pos), 0) + 1
                done = sorted(b for b, c in count.items() if c == n_tiles)
                rows = set()
                for b in done:
                    rows.update(range(b * rpb, min(M, (b + 1) * rpb)))
                assert rows == set(range(lo, hi)), (M, N, bn, cg, sms, rev, lo, hi)
                assert not any(blk(pos) in set(done) for pos in range(full, tiles))
                seen_overlap += bool(rem and done)
    assert seen_overlap >= 20
    assert lib.pgibbs_debug_tail_plan(0, 1280, 256, 2, 148, 0, out) != 0


def test_create_fails_loudly_without_gpu():
    lib = _lib.load()
    cfg = _lib.ModelConfig(arch=1, layers=1, embed_dim=64, heads=2, ffn_dim=128, vocab=33, max_positions=1024,
                           token_dropout=1, padding_idx=1, mask_idx=32, cls_idx=0, eos_idx=2)
    h = ctypes.c_void_p()
    assert lib.pgibbs_create(ctypes.byref(cfg), 0, ctypes.byref(h)) != 0
    assert b"CUDA" in lib.pgibbs_last_error() or b"device" in lib.pgibbs_last_error()
    with pytest.raises(_lib.EngineError):
        _lib.check(lib.pgibbs_op_sample(0, None, None, 1, 33, None, 20, 0, -1.0, None))


def test_algorithmic_flops_match_survey():
    import bench
    f = bench.algorithmic_flops_per_iter(get_config(bench.MODEL), 64, 258)
    assert abs(f / 1e12 - 22.20) < 0.01


def test_shard_range_covers_all_chains():
    for n in (1, 7, 64, 512):
        for w in (1, 2, 3, 8):
            spans = [parallel.shard_range(n, w, r) for r in range(w)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
            assert max(hi - lo for lo, hi in spans) - min(hi - lo for lo, hi in spans) <= 1


def _worker(rank, world, port, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    cfg = tiny_config("esm2", 1, 64, 2, 128)
    sd = synthetic_state_dict(cfg, 11) if rank == 0 else None
    got = parallel.broadcast_state_dict(sd, src=0, device="cpu")
    want = synthetic_state_dict(cfg, 11)
    same = all(torch.equal(got[k], want[k]) for k in want) and set(got) == set(want)
    # the packed blob: ONE broadcast, layout derived from the config on every rank, GEMM weights shipped as fp16
    from protein_gibbs_sampler_b200.weights import is_gemm_weight
    for arch, half in (("msa_transformer", True), ("esm1", True), ("roberta_large", False)):
        c2 = tiny_config(arch, 2, 64, 2, 128)
        full = synthetic_state_dict(c2, 5)
        blob = parallel.broadcast_weights(c2, full if rank == 0 else None, src=0, device="cpu", gemm_fp16=half)
        for k, v in full.items():
            if k == "lm_head.weight":
                continue
            w = v.half().float() if (half and is_gemm_weight(k)) else v
            same = same and torch.equal(blob[k], w)
        same = same and set(blob) == set(full) - {"lm_head.weight"}
    lo, hi = parallel.shard_range(5, world, rank)
    seqs = parallel.gather_sequences(["r%d_%d" % (rank, i) for i in range(lo, hi)])
    q.put((rank, same, seqs))
    dist.destroy_process_group()


def test_weight_broadcast_and_gather_gloo_world2():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + os.getpid() % 2000
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=120) for _ in procs)
    for p in procs:
        p.join(timeout=60)
    assert all(r[1] for r in res)
    assert res[0][2] == res[1][2] == ["r0_0", "r0_1", "r0_2", "r1_3", "r1_4"]
