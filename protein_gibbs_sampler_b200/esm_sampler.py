"""ESM_sampler: the reference's single-sequence Gibbs sampler API on the B200 engine.

Drop-in for /root/reference/src/pgen/esm_sampler.py: same class, method names, argument meaning, return
values and error messages.  What changes is where the loop body runs.  The reference executes lines 209-234
on the host (one forward, then one ``generate_step`` + host sync per residue); here the host only
  1. tokenises the seeds                                   (get_init_seq, :104-126)
  2. pre-draws the whole position schedule with Python's ``random`` in the reference's call order
     (get_random_target_index / get_target_index_in_order, :242-257) -- bit-identical masking/indexing,
  3. hands tokens + schedule to the engine, which runs all ``num_iters`` iterations on the GPU
     (mask scatter -> transformer forward -> top-k/categorical draw -> write-back) without a host sync,
  4. reads the tokens back and untokenises               (untokenize_batch, :84-93).

Residue draws: ``rng="device"`` (default) uses the engine's Philox generator seeded from torch's global
generator; ``rng="replay"`` pre-draws the Exp(1) variates ``Categorical.sample`` would consume from torch's
CPU generator, in the reference's order, so that identical logits give identical residues.
"""
import math
import random
import re
import time
from typing import Iterator, List, Tuple

import numpy as np
import torch
from tqdm import trange

from . import engine as _engine

ESM_ALLOWED_AMINO_ACIDS = "ACDEFGHIKLMNPQRSTVWY"


def _effective_k(top_k, n_valid, sample):
    return n_valid if (sample or top_k <= 0 or top_k > n_valid) else top_k


def generate_step(out, gen_idx, temperature=None, top_k=0, sample=False, valid_idx=None, device_id=0):
    """Draw one token id from ``out[gen_idx]`` (reference esm_sampler.py:8-45) with the engine's sampler
    kernel.  The Exp(1) variates come from torch's global CPU generator exactly as ``Categorical.sample``
    would draw them, so a seeded call reproduces the reference's choice.  Returns a 0-d LongTensor."""
    logits = torch.as_tensor(out)[gen_idx].detach().float().cpu().reshape(1, -1)
    if valid_idx is None:
        valid_idx = list(range(logits.shape[1]))
    k = _effective_k(top_k, len(valid_idx), sample)
    noise = torch.ones(1, len(valid_idx))
    noise[:, :k] = torch.empty(1, k).exponential_(1)
    tok = _engine.op_sample(logits, noise, valid_idx, top_k=k, temperature=temperature, device_id=device_id)
    return torch.tensor(int(tok[0]))


def parse_device(device):
    """Device-string grammar and errors of ESM_sampler.__init__ (reference esm_sampler.py:66-79)."""
    if device == "gpu":
        device = "cuda:0"
    use_cuda = False
    if re.match("^cuda:[0-9]+$", device):
        if not torch.cuda.is_available():
            raise Exception("gpu requested, but No Cuda devices found")
        use_cuda = True
        if int(device.split(":")[1]) >= torch.cuda.device_count():
            raise Exception("Invalid cuda device number: " + device)
    elif device != "cpu":
        raise Exception("Invalid device: " + device)
    return device, use_cuda


def in_order_targets(indexes, next_i, num_positions):
    """Cyclic cursor over ``indexes`` (reference :248-257): returns (last_i, picked positions)."""
    picked = []
    n = len(indexes)
    for _ in range(num_positions):
        next_i = (next_i + 1) % n
        picked.append(indexes[next_i])
    return next_i, picked


class SchedulePlan:
    """Positions for every iteration of one batch, laid out for pgibbs_set_schedule."""

    def __init__(self, positions, n_iters, P, iter_stride, chain_stride, has_duplicates):
        self.positions = np.ascontiguousarray(positions, dtype=np.int32).reshape(-1)
        self.n_iters, self.P = n_iters, P
        self.iter_stride, self.chain_stride = iter_stride, chain_stride
        self.has_duplicates = has_duplicates

    def targets(self, it, chain):
        o = it * self.iter_stride + chain * self.chain_stride
        return self.positions[o:o + self.P].tolist()

    def dense(self, n_chains):
        """Positions as an explicit [n_iters, n_chains, P] array (shared lists broadcast)."""
        if self.iter_stride == 0:
            p = self.positions[:self.P].reshape(1, 1, self.P)
        elif self.chain_stride == 0:
            p = self.positions.reshape(self.n_iters, 1, self.P)
        else:
            p = self.positions.reshape(self.n_iters, n_chains, self.P)
        return np.broadcast_to(p, (self.n_iters, n_chains, self.P))

    def slice_chains(self, lo, hi, n_chains):
        """The plan of chains [lo, hi) out of n_chains (a shared list -- strides 0 -- is everybody's)."""
        if self.chain_stride == 0:
            return self
        pos = self.positions.reshape(self.n_iters, n_chains, self.P)[:, lo:hi]
        return SchedulePlan(pos, self.n_iters, self.P, (hi - lo) * self.P, self.P, self.has_duplicates)


def draw_replay_noise(n_iters, rows, n_valid, top_k, burnin):
    """Exp(1) variates in the order the reference loop consumes them (iteration, chain, slot)."""
    ks = [_effective_k(top_k, n_valid, it < burnin) for it in range(n_iters)]
    stride = max(ks) if ks else 1
    noise = torch.ones(n_iters, rows, stride)
    for it, k in enumerate(ks):
        noise[it, :, :k] = torch.empty(rows, k).exponential_(1)
    return noise, stride


def strided_mask_plan(L, n_copies, start):
    """Positions masked by each strided copy: row i = start + (i, i+n, i+2n, ... < L), padded to P = ceil(L/n) slots
    by repeating the row's first position.  Returns (positions [n,P] int32, valid [n,P] bool)."""
    P = -(-L // n_copies)
    pos = np.empty((n_copies, P), dtype=np.int32)
    valid = np.zeros((n_copies, P), dtype=bool)
    for i in range(n_copies):
        p = np.arange(i, L, n_copies, dtype=np.int32) + start
        pos[i, :len(p)] = p
        pos[i, len(p):] = p[0]
        valid[i, :len(p)] = True
    return pos, valid


def score_strided(engine, unit_tokens, true_row, start, L, n_copies, batch_size, mask, row):
    """log p(true residue) at each of the L positions of one sequence (row=-1) or of row ``row`` of one MSA, from
    ``n_copies`` strided-masked copies of ``unit_tokens`` ([1,T] or [1,R,T]) scored ``batch_size`` copies per forward
    on the device.  Returns a list of L Python floats in position order."""
    true_row = np.asarray(true_row, dtype=np.int64)
    if mask:
        pos, valid = strided_mask_plan(L, n_copies, start)
    else:   # one unmasked copy, every position scored from it
        pos, valid = np.arange(start, start + L, dtype=np.int32)[None], np.ones((1, L), dtype=bool)
        n_copies = 1
    P = pos.shape[1]
    targets = np.where(valid, true_row[pos], -1).astype(np.int32)
    per_pos = np.zeros(L, dtype=np.float32)
    batch_size = max(1, int(batch_size or n_copies))
    for b0 in range(0, n_copies, batch_size):
        b1 = min(b0 + batch_size, n_copies)
        engine.set_tokens(unit_tokens.repeat(b1 - b0, *([1] * (unit_tokens.dim() - 1))))
        engine.set_schedule(pos[b0:b1], 1, P, (b1 - b0) * P, P, False)
        lp = engine.score(targets[b0:b1], mask=mask, row=row).numpy()
        v = valid[b0:b1]
        per_pos[pos[b0:b1][v] - start] = lp[v]
    return [float(x) for x in per_pos]


class ESM_sampler():
    """Gibbs sampler over single sequences; see module docstring."""

    def __init__(self, model, device="cpu", rng="device"):
        """model: object with ``model``, ``alphabet`` and ``batch_converter`` (see ``models.py``)."""
        if rng not in ("device", "replay"):
            raise ValueError("rng must be 'device' or 'replay'")
        self.model = model
        self.rng = rng
        self.model.model = self.model.model.eval()
        self.device, self.cuda = parse_device(device)
        self.model.model.to(self.device)
        self.valid_aa_idx = sorted(self.model.alphabet.get_idx(tok) for tok in ESM_ALLOWED_AMINO_ACIDS)
        self.last_timing = {}
        # Chains sharded across GPUs (parallel.shard_sampler): (rank, world_size, all_gather) or None.  Every rank
        # runs the same host code with the same RNG state, pre-draws the schedule (and replay noise) of ALL chains,
        # runs its contiguous slice and gathers the final tokens, so the result equals the single-GPU run's.
        self.shard = None

    # ------------------------------------------------------------------ host helpers (reference API)
    def untokenize_batch(self, batch, bos, eos):
        lo = 1 if bos else 0
        rows = batch.tolist() if isinstance(batch, torch.Tensor) else batch
        get_tok = self.model.alphabet.get_tok
        return ["".join(get_tok(t) for t in (row[lo:-1] if eos else row[lo:])) for row in rows]

    @staticmethod
    def clean_seed_seq(seed_to_clean):
        cleaned = seed_to_clean.upper()
        bad = set(cleaned) - set(ESM_ALLOWED_AMINO_ACIDS)
        if bad:
            raise Exception("Invalid input character: " + ",".join(bad))
        return cleaned

    def get_init_seq(self, seed_seq, max_len, batch_size=1):
        """Seeds -> token tensor, right-filled with <mask> up to max_len (reference :104-126)."""
        if isinstance(seed_seq, list):
            chosen = random.choices(seed_seq, k=batch_size)
            batch = [(str(i), self.clean_seed_seq(s) + "<mask>" * (max_len - len(s))) for i, s in enumerate(chosen)]
        elif isinstance(seed_seq, str):
            fill = "<mask>" * (max_len - len(seed_seq))
            cleaned = self.clean_seed_seq(seed_seq)
            batch = [(str(i), cleaned + fill) for i in range(batch_size)]
        else:
            raise Exception("seed sequence should either be a string or list")
        return self.model.batch_converter(batch)[2]

    def get_random_target_index(self, batch_size, indexes, num_positions):
        return [random.sample(indexes, num_positions) for _ in range(batch_size)]

    def get_target_index_in_order(self, batch_size, indexes, next_i, num_positions):
        last_i, picked = in_order_targets(indexes, next_i, num_positions)
        return last_i, [picked] * batch_size

    def mask_target_indexes(self, batch, target_indexes):
        mask_idx = self.model.alphabet.mask_idx
        for b, targets in enumerate(target_indexes):
            for kk in targets:
                batch[b][kk] = mask_idx

    def calculate_indexes(self, indexes, leader_length, max_len, rollover_from_start):
        """Candidate positions and the in-order cursor, including the reference's cursor quirk (:264-274)."""
        if indexes is not None:
            return indexes, -1
        indexes = range(1, max_len + 1)
        if rollover_from_start:
            return indexes, -1
        return indexes[leader_length:], leader_length - 1

    # ------------------------------------------------------------------ schedule
    def plan_positions(self, batch_size, indexes, last_i, num_positions, in_order, num_iters):
        """Pre-draw target positions for all iterations in the reference's RNG call order (:210-218)."""
        dup = len(set(indexes)) != len(indexes)
        if num_positions <= 0:
            return SchedulePlan(list(indexes), num_iters, len(indexes), 0, 0, dup), last_i
        if in_order:
            rows = []
            for _ in range(num_iters):
                last_i, picked = in_order_targets(indexes, last_i, num_positions)
                rows.append(picked)
            return SchedulePlan(rows, num_iters, num_positions, num_positions, 0, dup), last_i
        pos = [self.get_random_target_index(batch_size, indexes, num_positions) for _ in range(num_iters)]
        return SchedulePlan(pos, num_iters, num_positions, batch_size * num_positions, num_positions, dup), last_i

    # ------------------------------------------------------------------ generate
    def generate(self, n_samples, seed_seq, batch_size=1, in_order=False, max_len=None, leader_length=0,
                 leader_length_percent=None, top_k=0, temperature=None, num_iters=10, burnin=float('inf'),
                 mask=True, num_positions=0, num_positions_percent=None, indexes=None, rollover_from_start=False,
                 show_progress_bar=True):
        """Generate ``n_samples`` sequences; arguments as in the reference (esm_sampler.py:128-172)."""
        if isinstance(seed_seq, str):
            sequence_length = len(seed_seq)
        elif isinstance(seed_seq, list):
            sequence_length = max(len(seed) for seed in seed_seq)
        else:
            raise ValueError("Unknown seed sequence format, expecting str or list")
        alphabet = self.model.alphabet

        if max_len is None:
            max_len = sequence_length
        if num_positions_percent is not None:
            num_positions = int(max_len * (num_positions_percent / 100))
        num_positions = max(num_positions, 0)
        if leader_length_percent is not None:
            leader_length = int(max_len * (leader_length_percent / 100))
        leader_length = max(leader_length, 0)

        sequences = []
        n_batches = math.ceil(n_samples / batch_size)
        for batch_n in trange(n_batches, disable=(not show_progress_bar)):
            batch = self.get_init_seq(seed_seq, max_len, batch_size)
            indexes, last_i = self.calculate_indexes(indexes, leader_length, max_len, rollover_from_start)
            num_positions = min(num_positions, len(indexes))
            if num_iters > 0 and len(indexes) > 0:
                plan, last_i = self.plan_positions(batch_size, indexes, last_i, num_positions, in_order, num_iters)
                batch = self.run_plan(batch, plan, top_k, temperature, burnin, mask)[:, 0]
            keep = n_samples - len(sequences) if batch_n == n_batches - 1 else batch_size
            sequences += self.untokenize_batch(batch, alphabet.prepend_bos, alphabet.append_eos)[0:keep]
        return sequences

    def generate_many(self, seed_seqs, in_order=False, max_len=None, leader_length=0, leader_length_percent=None,
                      top_k=0, temperature=None, num_iters=10, burnin=float('inf'), mask=True, num_positions=0,
                      num_positions_percent=None, indexes=None, rollover_from_start=False, max_batch=256):
        """What ``[self.generate(1, s, batch_size=1, **kw)[0] for s in seed_seqs]`` returns -- the loop of
        pgen_esm_from_fasta.py:27-33 -- with the independent one-chain calls folded into one device batch per
        sequence length.  The host RNG draws of every call (position schedule from ``random``,
        replay noise from torch) are made call by call in that loop's order, and ``seed_seqs`` may be a generator
        that itself draws from ``random`` (``random.choice(seeds)``) between them, so in replay mode the result is
        identical to the sequential calls."""
        calls = []
        for seed in seed_seqs:
            if not isinstance(seed, str):
                raise ValueError("Unknown seed sequence format, expecting str or list")
            ml = len(seed) if max_len is None else max_len
            npos = int(ml * (num_positions_percent / 100)) if num_positions_percent is not None else num_positions
            lead = int(ml * (leader_length_percent / 100)) if leader_length_percent is not None else leader_length
            tokens = self.get_init_seq(seed, ml, 1)
            idx, last_i = self.calculate_indexes(indexes, max(lead, 0), ml, rollover_from_start)
            plan = noise = None
            if num_iters > 0 and len(idx) > 0:
                plan, _ = self.plan_positions(1, idx, last_i, min(max(npos, 0), len(idx)), in_order, num_iters)
                if self.rng == "replay":
                    noise = draw_replay_noise(plan.n_iters, plan.P, len(self.valid_aa_idx), top_k, burnin)
            calls.append((tokens, plan, noise))
        alphabet = self.model.alphabet
        final = [c[0] for c in calls]
        groups = {}
        for i, (tokens, plan, _) in enumerate(calls):
            if plan is not None:
                groups.setdefault((tokens.shape[-1], plan.P), []).append(i)
        for members in groups.values():
            for m0 in range(0, len(members), max_batch):
                part = members[m0:m0 + max_batch]
                plans = [calls[i][1] for i in part]
                pos = np.concatenate([p.dense(1) for p in plans], axis=1)
                merged = SchedulePlan(pos, plans[0].n_iters, plans[0].P, len(part) * plans[0].P, plans[0].P,
                                      any(p.has_duplicates for p in plans))
                noise = None
                if self.rng == "replay":
                    noise = (torch.cat([calls[i][2][0] for i in part], dim=1), calls[part[0]][2][1])
                out = self.run_plan(torch.cat([calls[i][0] for i in part], dim=0), merged, top_k, temperature, burnin,
                                    mask, noise=noise)[:, 0]
                for j, i in enumerate(part):
                    final[i] = out[j:j + 1]
        return [self.untokenize_batch(t, alphabet.prepend_bos, alphabet.append_eos)[0] for t in final]

    def run_plan(self, tokens, plan, top_k, temperature, burnin, mask, noise=None):
        """Ship tokens + schedule to the GPU, run every iteration there, return the final tokens [B,R,T].
        ``noise``: (Exp(1) draws [n_iters, rows, stride], stride) already drawn by the caller (replay mode).
        Raises if the model has no CUDA engine: there is no CPU path."""
        t0 = time.perf_counter()
        engine = self.model.model.require_engine()
        n_chains = int(np.prod(tokens.shape[:-1]))
        stride = None
        if noise is not None:
            noise, stride = noise
        elif self.rng == "replay":   # drawn for ALL chains, in the reference's order, whatever the sharding
            noise, stride = draw_replay_noise(plan.n_iters, n_chains * plan.P, len(self.valid_aa_idx), top_k, burnin)
        else:
            device_seed = int(torch.randint(0, 2 ** 62, (1,)).item())
        lo, hi = 0, n_chains
        if self.shard is not None:
            from .parallel import shard_range
            rank, world, _ = self.shard
            lo, hi = shard_range(n_chains, world, rank)
            if noise is not None:
                noise = noise.reshape(plan.n_iters, n_chains, plan.P, stride)[:, lo:hi].reshape(plan.n_iters, -1, stride)
            plan, tokens = plan.slice_chains(lo, hi, n_chains), tokens[lo:hi]
        if hi > lo:
            engine.set_tokens(tokens)
            engine.set_schedule(plan.positions, plan.n_iters, plan.P, plan.iter_stride, plan.chain_stride,
                                plan.has_duplicates)
            if noise is not None:
                engine.set_noise(noise.contiguous(), stride)
            else:
                engine.set_noise(None)
                engine.set_device_rng(device_seed)
            engine.set_chain_offset(lo)
        t1 = time.perf_counter()
        if hi > lo:
            engine.run(0, plan.n_iters, burnin, top_k, temperature, mask, self.valid_aa_idx)
        t2 = time.perf_counter()
        out = engine.get_tokens() if hi > lo else torch.empty((0, 1, tokens.shape[-1]), dtype=torch.int64)
        if self.shard is not None:
            out = torch.cat(self.shard[2](out), dim=0)   # every rank ends up with all chains, in chain order
        t3 = time.perf_counter()
        # host-side trace of the last batch (the reference has no tracing; this is the engine's)
        self.last_timing = {"upload_s": t1 - t0, "enqueue_s": t2 - t1, "wait_and_download_s": t3 - t2}
        return out

    # ------------------------------------------------------------------ scoring (shares the forward)
    def log_likelihood(self, seq, with_masking=True, verbose=False, mask_distance=float("inf"),
                       batch_size=None) -> Tuple[float, List[float]]:
        return next(self.log_likelihood_batch([seq], with_masking, verbose, mask_distance, batch_size))

    def log_likelihood_batch(self, seq_list, with_masking=True, verbose=False, mask_distance=float("inf"),
                             batch_size=None) -> Iterator[Tuple[float, List[float]]]:
        """Pseudo-log-likelihood with strided masking (reference :288-363): copy i of the sequence masks positions
        i, i+n, ... (n = min(mask_distance, L)); each position is scored from the copy that masks it.  With the B200
        engine the copies are masked on the device, the LM head runs on the masked rows only and the log_softmax +
        gather at the true residue is fused into it (``Engine.score``): L floats come back instead of n*T*V logits.
        Any other ``model.model`` (the downward duck-typed interface) is called for logits as the reference does.
        Every forward here is over equal-length rows, so no <pad> reaches the engine."""
        alphabet = self.model.alphabet
        if batch_size is None:
            batch_size = len(seq_list)
        start = 1 if alphabet.prepend_bos else 0
        on_engine = hasattr(self.model.model, "require_engine")
        for seq in seq_list:
            cleaned = self.clean_seed_seq(seq)
            L = len(cleaned)
            true_toks = self.model.batch_converter([("0", cleaned)])[2][0]
            n_copies = int(min(mask_distance, L)) if with_masking else 1
            if on_engine:
                per_pos = score_strided(self.model.model.require_engine(), true_toks[None], true_toks, start, L,
                                        n_copies, batch_size, with_masking, row=-1)
            elif with_masking:
                toks = true_toks.repeat(n_copies, 1)
                for i in range(n_copies):
                    toks[i, start + i:start + L:n_copies] = alphabet.mask_idx
                per_pos = [None] * L
                for b0 in range(0, n_copies, batch_size):
                    lp = torch.log_softmax(self.model.model(toks[b0:b0 + batch_size])["logits"], dim=-1)
                    for j in range(lp.shape[0]):
                        i = b0 + j
                        for pos in range(i, L, n_copies):
                            per_pos[pos] = lp[j, start + pos, true_toks[start + pos]].item()
            else:
                lp = torch.log_softmax(self.model.model(true_toks[None])["logits"], dim=-1)[0]
                per_pos = [lp[start + pos, true_toks[start + pos]].item() for pos in range(L)]
            # the reference accumulates copy by copy; keep its output ordering
            ordered = [per_pos[pos] for i in range(n_copies) for pos in range(i, L, n_copies)]
            total = np.float32(0.0)   # the reference sums 0-d float32 tensors
            for v in ordered:
                total = np.float32(total + np.float32(v))
            yield (float(total / np.float32(L)), ordered)
