"""C-ABI contract of the engine on a GPU: error behaviour, determinism, iteration ranges, duplicate positions.
Complements tests/test_gpu.py (numerical parity) -- everything here goes through libpgibbs.so."""
import random

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def sampler(gpu_lib):
    from protein_gibbs_sampler_b200 import models
    from protein_gibbs_sampler_b200.config import tiny_config
    from protein_gibbs_sampler_b200.esm_sampler import ESM_sampler
    assert torch.cuda.is_available()
    cfg = tiny_config("esm2", 2, 128, 2, 256)
    return ESM_sampler(models.CustomModel(cfg, seed=4), device="cuda:0", rng="replay")


def _tokens(sampler, B=3, L=20):
    return sampler.get_init_seq("MKTAYIAKQRQISFVKSHFS"[:L], L, B)


def test_rejected_inputs_raise_with_a_message(sampler):
    from protein_gibbs_sampler_b200._lib import EngineError
    eng = sampler.model.model.engine
    tok = _tokens(sampler)
    bad = tok.clone(); bad[1, 5] = 1                      # <pad>
    with pytest.raises(EngineError, match="pad"):
        eng.set_tokens(bad)
    bad = tok.clone(); bad[0, 2] = 33                     # outside the 33-token vocabulary
    with pytest.raises(EngineError, match="vocabulary"):
        eng.set_tokens(bad)
    eng.set_tokens(tok)
    with pytest.raises(EngineError, match="outside"):
        eng.set_schedule(np.array([1, 2, 22], dtype=np.int32), 1, 3, 3, 0)   # position 22 >= T = 22
    with pytest.raises(EngineError, match="exceeds"):
        eng.set_schedule(np.arange(23, dtype=np.int32) % 22, 1, 23, 23, 0)   # P > T
    eng.set_schedule(np.array([1, 2, 3], dtype=np.int32), 1, 3, 3, 0)
    with pytest.raises(EngineError, match="outside the schedule"):
        eng.run(0, 2, 0, 0, None, True, sampler.valid_aa_idx)               # 2 iterations, schedule has 1
    with pytest.raises(EngineError, match="MSA"):
        eng.run_single(0, 1, 0, 1, None, -1, 0, sampler.valid_aa_idx)       # not an MSA model
    with pytest.raises(EngineError, match="n_valid"):
        eng.run(0, 1, 0, 0, None, True, [])


def test_same_seed_same_sequences_and_iteration_ranges_compose(sampler):
    """Deterministic kernels: one `generate` repeated gives identical strings; running iterations [0,4) in one call
    equals [0,1) + [1,3) + [3,4) (no host state between iterations other than the schedule / noise slices)."""
    kw = dict(seed_seq="MKTAYIAKQRQISFVKSHFS", batch_size=4, num_iters=4, top_k=4, burnin=2, num_positions=6,
              show_progress_bar=False)
    outs = []
    for _ in range(2):
        random.seed(9); torch.manual_seed(9)
        outs.append(sampler.generate(4, **kw))
    assert outs[0] == outs[1] and len(set(outs[0])) > 1
    eng = sampler.model.model.engine
    tok = _tokens(sampler, 4)
    idx, last = sampler.calculate_indexes(None, 0, 20, False)
    random.seed(2)
    plan, _ = sampler.plan_positions(4, idx, last, 6, False, 4)
    from protein_gibbs_sampler_b200.esm_sampler import draw_replay_noise
    torch.manual_seed(2)
    noise, stride = draw_replay_noise(4, 4 * 6, 20, 4, 2)
    results = []
    for pieces in ([(0, 4)], [(0, 1), (1, 2), (3, 1)]):
        eng.set_tokens(tok)
        eng.set_schedule(plan.positions, plan.n_iters, plan.P, plan.iter_stride, plan.chain_stride)
        eng.set_noise(noise, stride)
        for first, n in pieces:
            eng.run(first, n, 2, 4, None, True, sampler.valid_aa_idx)
        results.append(eng.get_tokens())
    assert torch.equal(results[0], results[1])


def test_duplicate_positions_last_slot_wins_like_the_reference_loop(sampler):
    """User `indexes` with repeats: the reference writes slot after slot (esm_sampler.py:225-234), so the LAST slot of a
    position decides; all slots see the same logits but consume different noise."""
    from oracle.fair_esm import OracleModel
    from oracle.gibbs_loop import esm_generate
    from protein_gibbs_sampler_b200.weights import synthetic_state_dict
    cfg = sampler.model.cfg
    om = OracleModel(cfg, synthetic_state_dict(cfg, 4))
    kw = dict(seed_seq="MKTAYIAKQRQISFVKSHFS", batch_size=3, num_iters=2, burnin=5, indexes=[3, 7, 3, 9, 7, 7])
    random.seed(1); torch.manual_seed(1)
    want = esm_generate(om, 3, **kw)
    random.seed(1); torch.manual_seed(1)
    got = sampler.generate(3, show_progress_bar=False, **kw)
    diff = sum(a != b for x, y in zip(got, want) for a, b in zip(x, y))
    assert diff <= 1, (got, want)
    seed = kw["seed_seq"]
    assert all(g[i] == seed[i] for g in got for i in range(20) if i + 1 not in (3, 7, 9))


def test_device_rng_mode_is_seeded_by_torch(gpu_lib):
    """rng='device' (the default): the Philox key comes from torch's global generator, so torch.manual_seed makes a run
    repeatable and different seeds give different chains."""
    from protein_gibbs_sampler_b200 import models
    from protein_gibbs_sampler_b200.config import tiny_config
    from protein_gibbs_sampler_b200.esm_sampler import ESM_sampler
    s = ESM_sampler(models.CustomModel(tiny_config("esm2", 2, 128, 2, 256), seed=4), device="cuda:0")
    kw = dict(seed_seq="MKTAYIAKQRQISFVKSHFS", batch_size=4, num_iters=3, show_progress_bar=False)
    runs = []
    for sd in (5, 5, 6):
        random.seed(0); torch.manual_seed(sd)
        runs.append(s.generate(4, **kw))
    assert runs[0] == runs[1] and runs[0] != runs[2]
