"""Chains sharded across GPUs: every rank pre-draws the whole schedule / noise with the same RNG
state, runs its contiguous slice, and the tokens are gathered once -- the result must equal the single-GPU run's.

CPU: host logic with a stand-in engine (a deterministic function of schedule, noise / device seed and GLOBAL row index),
ranks simulated in sequence and as two real gloo processes.  GPU: the real engine, two "ranks" run one after the other
on cuda:0, in replay and in device-RNG mode."""
import multiprocessing as mp
import os
import random

import numpy as np
import pytest
import torch

from protein_gibbs_sampler_b200 import parallel
from protein_gibbs_sampler_b200.alphabet import Alphabet
from protein_gibbs_sampler_b200.esm_msa_sampler import ESM_MSA_sampler
from protein_gibbs_sampler_b200.esm_sampler import ESM_sampler

SEED_SEQ = "MKTAYIAKQRQISFVKSHFS"
MSA = ["MKTAYIAK-RQ", "MKSAY-AKQRQ", "MRTAYIAKQ-Q"]


class FakeEngine:
    """Same call sequence as engine.Engine; the 'model' writes valid[(f(noise or seed, global row, iter)) % n_valid]."""

    def set_tokens(self, t):
        t = torch.as_tensor(t)
        self.tok = (t[:, None, :] if t.dim() == 2 else t).clone().to(torch.int64)

    def set_schedule(self, positions, n_iters, P, iter_stride, chain_stride, has_duplicates=False):
        self.sched = (np.asarray(positions, dtype=np.int64).reshape(-1), n_iters, P, iter_stride, chain_stride)

    def set_noise(self, noise, stride=0):
        self.noise = None if noise is None else torch.as_tensor(noise).reshape(self.sched[1], -1, stride)

    def set_device_rng(self, seed):
        self.seed = seed

    def set_chain_offset(self, first):
        self.first = first

    def run(self, first_iter, num_iters, burnin, top_k, temperature, mask, valid):
        pos, n_iters, P, istr, cstr = self.sched
        flat = self.tok.view(-1, self.tok.shape[-1])
        for it in range(first_iter, first_iter + num_iters):
            for chain in range(flat.shape[0]):
                for s in range(P):
                    row = chain * P + s
                    if self.noise is not None:
                        key = int(self.noise[it, row, 0].item() * 1e6)
                    else:
                        key = self.seed + 7919 * (self.first * P + row) + 104729 * it
                    flat[chain, pos[it * istr + chain * cstr + s]] = valid[key % len(valid)]

    def get_tokens(self):
        return self.tok.clone()


class FakeModule:
    def __init__(self):
        self.engine = FakeEngine()

    def eval(self):
        return self

    def to(self, device):
        return self

    def require_engine(self):
        return self.engine


class FakeModel:
    def __init__(self, msa=False):
        self.alphabet = Alphabet.msa() if msa else Alphabet.esm1b()
        self.batch_converter = self.alphabet.get_batch_converter()
        self.model = FakeModule()


def _own_slice_gather(rank, world, n_total):
    """all_gather stand-in for ranks simulated one after the other: this rank's chains in place, zeros elsewhere."""
    def gather(local):
        parts = []
        for r in range(world):
            lo, hi = parallel.shard_range(n_total, world, r)
            parts.append(local if r == rank else torch.zeros((hi - lo,) + tuple(local.shape[1:]), dtype=local.dtype))
        return parts
    return gather


def _run(sampler_factory, kwargs, seed, shard=None):
    s = sampler_factory()
    if shard is not None:
        parallel.shard_sampler(s, *shard)
    random.seed(seed)
    torch.manual_seed(seed)
    return s.generate(**kwargs)


@pytest.mark.parametrize("rng", ["replay", "device"])
@pytest.mark.parametrize("world", [2, 3, 5])
def test_sharded_generate_equals_unsharded_host_logic(rng, world):
    kw = dict(n_samples=7, seed_seq=[SEED_SEQ, SEED_SEQ[:12]], batch_size=4, max_len=20, num_iters=3, top_k=3, burnin=1,
              num_positions=5, show_progress_bar=False)

    def make():
        return ESM_sampler(FakeModel(), device="cpu", rng=rng)
    want = _run(make, kw, 3)
    assert len(want) == 7 and len(set(want)) > 1
    for rank in range(world):
        got = _run(make, kw, 3, (rank, world, _own_slice_gather(rank, world, 4)))
        for b0 in (0, 4):                                   # two outer batches of 4 chains (the last one truncated)
            lo, hi = parallel.shard_range(4, world, rank)
            assert got[b0 + lo:min(b0 + hi, 7)] == want[b0 + lo:min(b0 + hi, 7)], (rank, world)
    # in-order and all-positions schedules are one shared list (strides 0): the slice is the plan itself
    kw2 = dict(kw, in_order=True, num_positions=4)
    want2 = _run(make, kw2, 4)
    got2 = _run(make, kw2, 4, (1, 2, _own_slice_gather(1, 2, 4)))
    assert got2[2:4] == want2[2:4]


@pytest.mark.parametrize("rng", ["replay", "device"])
def test_sharded_msa_generate_splits_whole_msas(rng):
    kw = dict(n_samples=9, seed_msa=MSA, batch_size=3, num_iters=2, top_k=2, burnin=1, num_positions=3,
              show_progress_bar=False)

    def make():
        return ESM_MSA_sampler(FakeModel(msa=True), device="cpu", rng=rng)
    want = _run(make, kw, 5)
    for rank in range(2):
        got = _run(make, kw, 5, (rank, 2, _own_slice_gather(rank, 2, 3)))
        lo, hi = parallel.shard_range(3, 2, rank)               # MSAs, each of 3 rows
        assert got[lo * 3:hi * 3] == want[lo * 3:hi * 3]


def _gloo_worker(rank, world, port, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    import torch.distributed as dist
    dist.init_process_group("gloo", rank=rank, world_size=world)
    kw = dict(n_samples=5, seed_seq=SEED_SEQ, batch_size=5, num_iters=3, top_k=2, burnin=1, num_positions=6,
              show_progress_bar=False)

    def make():
        return ESM_sampler(FakeModel(), device="cpu", rng="replay")
    want = _run(make, kw, 8)
    s = parallel.shard_sampler(make())        # rank / world from the process group, real all_gather
    random.seed(8)
    torch.manual_seed(8)
    got = s.generate(**kw)
    q.put((rank, got == want, got))
    dist.destroy_process_group()


def test_sharded_generate_gloo_world2():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 31500 + os.getpid() % 2000
    procs = [ctx.Process(target=_gloo_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=180) for _ in procs)
    for p in procs:
        p.join(timeout=60)
    assert all(r[1] for r in res) and res[0][2] == res[1][2]


@pytest.mark.gpu
@pytest.mark.parametrize("rng", ["replay", "device"])
def test_sharded_generate_on_engine_equals_single_gpu(rng, gpu_lib):
    from protein_gibbs_sampler_b200 import models
    from protein_gibbs_sampler_b200.config import tiny_config
    cfg = tiny_config("esm2", 2, 128, 2, 256)
    model = models.CustomModel(cfg, seed=6)
    kw = dict(n_samples=6, seed_seq=SEED_SEQ, batch_size=6, num_iters=3, top_k=3, burnin=1, num_positions=5,
              show_progress_bar=False)

    def make():
        return ESM_sampler(model, device="cuda:0", rng=rng)
    want = _run(make, kw, 12)
    assert len(set(want)) > 1
    for rank in range(2):
        got = _run(make, kw, 12, (rank, 2, _own_slice_gather(rank, 2, 6)))
        lo, hi = parallel.shard_range(6, 2, rank)
        assert got[lo:hi] == want[lo:hi], (rng, rank)


def test_generate_many_folds_one_chain_calls_in_rng_order():
    """pgen_esm_from_fasta's loop (reference :27-33) folded into device batches: the seed choice, the position
    schedule and the replay noise of every one-chain call are drawn in the loop's order, the chains run grouped by
    sequence length -- same output as the call-by-call loop (stand-in engine keyed by the noise it is handed)."""
    from protein_gibbs_sampler_b200.fasta import unalign
    seeds = ["MKTAYIAKQR-ISFVK", "MKT.AYLAKQRQISFVKSH", "mrtayiakqrq-sfvk", "MKTAYIAKQRQISFVKSHFSRQ"]
    for kw in (dict(num_iters=4, num_positions=3, burnin=2, top_k=2), dict(num_iters=3, in_order=True, num_positions=2),
               dict(num_iters=2, num_positions_percent=50, leader_length=2), dict(num_iters=0), dict(num_iters=3)):
        s = ESM_sampler(FakeModel(), device="cpu", rng="replay")
        random.seed(9); torch.manual_seed(9)
        folded = s.generate_many((unalign(random.choice(seeds))[0] for _ in range(9)), max_batch=4, **kw)
        random.seed(9); torch.manual_seed(9)
        looped = [s.generate(1, unalign(random.choice(seeds))[0], batch_size=1, show_progress_bar=False, **kw)[0]
                  for _ in range(9)]
        assert folded == looped, kw
    with pytest.raises(ValueError, match="expecting str"):
        s.generate_many([["MKT"]])


class ScoringEngine(FakeEngine):
    """`Engine.score` stand-in: log p of slot (chain, j) = -(position + target / 100), or 0 for a padding slot; also
    checks what the device path relies on -- unmasked tokens resident, targets = the true tokens at the positions."""

    def score(self, targets, mask=True, row=-1):
        pos, n_iters, P, istr, cstr = self.sched
        t = np.asarray(targets).reshape(-1, P)
        rows = self.tok.shape[1]
        out = np.zeros(t.shape, dtype=np.float32)
        for c in range(t.shape[0]):
            for j in range(P):
                p_ = int(pos[c * cstr + j])
                seq = self.tok[c, row if row >= 0 else 0] if rows > 1 or row < 0 else self.tok[c, 0]
                if t[c, j] >= 0:
                    assert int(seq[p_]) == int(t[c, j])
                    out[c, j] = -(p_ + t[c, j] / 100.0)
        self.calls = getattr(self, "calls", 0) + 1
        return torch.from_numpy(out)


@pytest.mark.parametrize("L,mask_distance,batch_size,with_masking", [(11, float("inf"), None, True), (11, 4, 2, True),
                                                                    (11, 4, None, True), (7, 3, 1, True),
                                                                    (9, float("inf"), None, False)])
def test_device_scoring_host_mapping(L, mask_distance, batch_size, with_masking):
    """log_likelihood_batch on the engine path: schedule / targets handed to `score`, and the way its [copies, P]
    result is put back into the reference's copy-major output order (esm_sampler.py:331-352)."""
    m = FakeModel()
    m.model.engine = ScoringEngine()
    s = ESM_sampler(m, device="cpu")
    seq = (SEED_SEQ * 2)[:L]
    toks = m.batch_converter([("0", seq)])[2][0].tolist()
    mean, each = next(s.log_likelihood_batch([seq], with_masking=with_masking, mask_distance=mask_distance,
                                             batch_size=batch_size))
    n = int(min(mask_distance, L)) if with_masking else 1
    order = [p for i in range(n) for p in range(i, L, n)]
    assert each == pytest.approx([-((1 + p) + toks[1 + p] / 100.0) for p in order], abs=1e-5)
    assert mean == pytest.approx(sum(each) / L, abs=1e-4)
    # batch_size=None means len(seq_list) masked copies per forward, as in the reference (esm_sampler.py:305-306)
    assert m.model.engine.calls == -(-n // (batch_size or 1))


def test_device_scoring_host_mapping_msa():
    m = FakeModel(msa=True)
    m.model.engine = ScoringEngine()
    s = ESM_MSA_sampler(m, device="cpu")
    toks = m.batch_converter([[(str(i), q) for i, q in enumerate(MSA)]])[2][0]
    for target, count_gaps in ((0, False), (1, False), (2, True)):
        mean, each = s.log_likelihood(MSA, target_index=target, mask_distance=3, count_gaps=count_gaps)
        L = len(MSA[0])
        order = [p for i in range(3) for p in range(i, L, 3) if count_gaps or MSA[target][p] != "-"]
        assert each == pytest.approx([-((1 + p) + int(toks[target, 1 + p]) / 100.0) for p in order], abs=1e-5)
