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
