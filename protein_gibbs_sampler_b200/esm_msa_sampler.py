"""ESM_MSA_sampler: the reference's MSA Gibbs sampler API on the B200 engine.

Drop-in for /root/reference/src/pgen/esm_msa_sampler.py (class, methods, arguments, return values, error
messages).  As in ``esm_sampler.py`` the host tokenises, pre-draws the position schedule with Python's
``random`` in the reference's call order -- one ``random.sample`` per (MSA, row) per iteration (:272-279) --
and the engine runs the loop body (:221-248, or :126-145 for ``generate_single``) on the GPU.
"""
import math
import random
from typing import Iterator, List, Tuple

import numpy as np
import torch
from tqdm import trange

from .esm_sampler import (SchedulePlan, draw_replay_noise, generate_step, in_order_targets, parse_device,  # noqa: F401
                          score_strided)

ESM_MSA_ALLOWED_AMINO_ACIDS = "-ACDEFGHIKLMNPQRSTVWY"
ESM_MSA_GAP_CHARACTERS = "-"


def partition(input_list, num_partitions):
    """Split into ``num_partitions`` consecutive bins; the remainder goes one-each to the first bins
    (reference esm_msa_sampler.py:13-31, pinned by test_esm_msa_sampler.py:538-557)."""
    n = len(input_list)
    num_partitions = min(num_partitions, n)
    if num_partitions <= 0:
        return []
    base, extra = divmod(n, num_partitions)
    out, lo = [], 0
    for i in range(num_partitions):
        hi = lo + base + (1 if i < extra else 0)
        out.append(list(input_list[lo:hi]))
        lo = hi
    return out


class ESM_MSA_sampler():
    def __init__(self, model, device="cpu", rng="device"):
        if rng not in ("device", "replay"):
            raise ValueError("rng must be 'device' or 'replay'")
        self.model = model
        self.rng = rng
        self.model.model = self.model.model.eval()
        self.device, self.cuda = parse_device(device)
        self.model.model.to(self.device)
        self.valid_aa_idx = sorted(self.model.alphabet.get_idx(tok) for tok in ESM_MSA_ALLOWED_AMINO_ACIDS)
        self.shard = None   # see ESM_sampler.shard / parallel.shard_sampler
        self.toks = [self.model.alphabet.get_tok(idx) for idx in self.valid_aa_idx]

    # ------------------------------------------------------------------ host helpers (reference API)
    def untokenize_batch(self, batch):
        """All rows of all MSAs, flattened in (msa, row) order, <cls> column dropped (:68-75)."""
        msas = batch.tolist() if isinstance(batch, torch.Tensor) else batch
        get_tok = self.model.alphabet.get_tok
        return ["".join(get_tok(t) for t in row[1:]) for msa in msas for row in msa]

    def clean_seed_seq(self, seq):
        seq = seq.upper()
        bad = set(seq) - set(ESM_MSA_ALLOWED_AMINO_ACIDS)
        if bad:
            raise Exception("Invalid input character: " + ",".join(bad))
        return seq

    def get_init_msa(self, seed_msa, max_len, batch_size=1):
        padded = []
        for i, seq in enumerate(seed_msa):
            seq = self.clean_seed_seq(seq)
            padded.append((str(i), seq + "<mask>" * (max_len - len(seq))))
        return self.model.batch_converter([padded] * batch_size)[2]

    def mask_target_indexes(self, batch, target_indexes):
        mask_idx = self.model.alphabet.mask_idx
        for b in range(len(batch)):
            for r in range(len(batch[b])):
                for kk in target_indexes[b][r]:
                    batch[b][r][kk] = mask_idx

    def mask_target_indexes_single(self, batch, target_indexes, seq_index):
        mask_idx = self.model.alphabet.mask_idx
        for b in range(len(batch)):
            for kk in target_indexes:
                batch[b][seq_index][kk] = mask_idx

    def get_target_indexes_all_positions(self, batch_size, indexes, num_sequences):
        return [[indexes] * num_sequences for _ in range(batch_size)]

    def get_random_target_index(self, batch_size, indexes, num_positions, num_sequences):
        return [[random.sample(indexes, num_positions) for _ in range(num_sequences)] for _ in range(batch_size)]

    def get_target_index_in_order(self, batch_size, indexes, next_i, num_positions, num_sequences):
        last_i, picked = in_order_targets(indexes, next_i, num_positions)
        return last_i, [[picked] * num_sequences for _ in range(batch_size)]

    def calculate_indexes(self, indexes, leader_length, max_len, rollover_from_start):
        if indexes is not None:
            return indexes, -1
        indexes = list(range(1, max_len + 1))
        if rollover_from_start:
            return indexes, -1
        return indexes[leader_length:], leader_length - 1

    # ------------------------------------------------------------------ engine plumbing
    def _noise(self, engine, n_iters, rows, top_k, burnin):
        if self.rng == "replay":
            noise, stride = draw_replay_noise(n_iters, rows, len(self.valid_aa_idx), top_k, burnin)
            engine.set_noise(noise, stride)
        else:
            engine.set_noise(None)
            engine.set_device_rng(int(torch.randint(0, 2 ** 62, (1,)).item()))

    def plan_positions(self, batch_size, num_sequences, indexes, last_i, num_positions, in_order, num_iters):
        dup = len(set(indexes)) != len(indexes)
        if num_positions <= 0:
            return SchedulePlan(list(indexes), num_iters, len(indexes), 0, 0, dup), last_i
        if in_order:
            rows = []
            for _ in range(num_iters):
                last_i, picked = in_order_targets(indexes, last_i, num_positions)
                rows.append(picked)
            return SchedulePlan(rows, num_iters, num_positions, num_positions, 0, dup), last_i
        pos = [self.get_random_target_index(batch_size, indexes, num_positions, num_sequences)
               for _ in range(num_iters)]
        n_chains = batch_size * num_sequences
        return SchedulePlan(pos, num_iters, num_positions, n_chains * num_positions, num_positions, dup), last_i

    def run_plan(self, tokens, plan, top_k, temperature, burnin, mask):
        """Run every iteration of one round on the GPU; returns the final tokens [B,R,C].  No CPU path.
        With ``self.shard`` set (parallel.shard_sampler) whole MSAs are split across ranks: schedule and replay noise
        are drawn for all of them, this rank runs MSAs [lo, hi) and the final tokens are all-gathered."""
        engine = self.model.model.require_engine()
        B, R = tokens.shape[0], tokens.shape[1]
        n_chains = B * R
        noise = stride = None
        if self.rng == "replay":
            noise, stride = draw_replay_noise(plan.n_iters, n_chains * plan.P, len(self.valid_aa_idx), top_k, burnin)
        else:
            device_seed = int(torch.randint(0, 2 ** 62, (1,)).item())
        lo, hi = 0, B
        if self.shard is not None:
            from .parallel import shard_range
            rank, world, _ = self.shard
            lo, hi = shard_range(B, world, rank)   # never split one MSA: its rows are coupled by the attention
            if noise is not None:
                noise = noise.reshape(plan.n_iters, n_chains, plan.P, stride)[:, lo * R:hi * R]
                noise = noise.reshape(plan.n_iters, -1, stride)
            plan, tokens = plan.slice_chains(lo * R, hi * R, n_chains), tokens[lo:hi]
        if hi > lo:
            engine.set_tokens(tokens)
            engine.set_schedule(plan.positions, plan.n_iters, plan.P, plan.iter_stride, plan.chain_stride,
                                plan.has_duplicates)
            if noise is not None:
                engine.set_noise(noise.contiguous(), stride)
            else:
                engine.set_noise(None)
                engine.set_device_rng(device_seed)
            engine.set_chain_offset(lo * R)
            engine.run(0, plan.n_iters, burnin, top_k, temperature, mask, self.valid_aa_idx)
            out = engine.get_tokens()
        else:
            out = torch.empty((0, R, tokens.shape[-1]), dtype=torch.int64)
        if self.shard is not None:
            out = torch.cat(self.shard[2](out), dim=0)
        return out

    # ------------------------------------------------------------------ generate
    def generate(self, n_samples, seed_msa, batch_size=1, in_order=False, max_len=None, leader_length=0,
                 leader_length_percent=None, top_k=0, temperature=None, num_iters=10, burnin=float('inf'),
                 mask=True, num_positions=0, num_positions_percent=None, indexes=None, rollover_from_start=False,
                 show_progress_bar=True):
        """Resample every row of ``batch_size`` copies of the seed MSA (reference :151-253)."""
        num_sequences = len(seed_msa)
        sequence_length = len(seed_msa[0])
        n_rounds = math.ceil(n_samples / num_sequences / batch_size)
        if num_positions_percent is not None:
            num_positions = int(sequence_length * (num_positions_percent / 100))
        num_positions = max(num_positions, 0)
        if leader_length_percent is not None:
            leader_length = int(sequence_length * (leader_length_percent / 100))
        leader_length = max(leader_length, 0)
        if max_len is None:
            max_len = sequence_length

        sequences = []
        for rnd in trange(n_rounds, disable=(not show_progress_bar)):
            batch = self.get_init_msa(seed_msa, max_len, batch_size)
            indexes, last_i = self.calculate_indexes(indexes, leader_length, max_len, rollover_from_start)
            num_positions = min(num_positions, len(indexes))
            if num_iters > 0 and len(indexes) > 0:
                plan, last_i = self.plan_positions(batch_size, num_sequences, indexes, last_i, num_positions,
                                                   in_order, num_iters)
                batch = self.run_plan(batch, plan, top_k, temperature, burnin, mask)
            out = self.untokenize_batch(batch)
            sequences += out[0:n_samples - len(sequences)] if rnd == n_rounds - 1 else out
        return sequences

    def generate_single(self, seed_msa, steps=10, passes=3, burn_in=1, target_index=0, k=1, exclude_positions=None):
        """One new sequence for row ``target_index`` (reference :101-147).  Each pass shuffles the positions,
        splits them into ``steps`` bins; per bin the LAST row is masked (reference quirk, :133) and row
        ``target_index`` is resampled at the bin's positions."""
        engine = self.model.model.require_engine()
        excluded = {i + 1 for i in (exclude_positions or [])}
        sequence_length = len(seed_msa[0])
        positions = [x for x in range(1, sequence_length + 1) if x not in excluded]
        batch = self.get_init_msa(seed_msa, sequence_length, 1)
        engine.set_tokens(batch)
        for pass_num in range(passes):
            random.shuffle(positions)
            bins = partition(positions, steps)
            i = 0
            while i < len(bins):  # consecutive bins of equal size share one schedule upload
                j = i
                while j < len(bins) and len(bins[j]) == len(bins[i]):
                    j += 1
                group = bins[i:j]
                P = len(group[0])
                engine.set_schedule([p for b in group for p in b], len(group), P, P, 0, False)
                burnin = float("inf") if pass_num < burn_in else 0
                self._noise(engine, len(group), P, k, burnin)
                engine.run_single(0, len(group), burnin, k, None, -1, target_index, self.valid_aa_idx)
                i = j
        return self.untokenize_batch(engine.get_tokens())[target_index]

    def generate_single_batch(self, seed_msa, n, steps=10, passes=3, burn_in=1, target_index=0, k=1,
                              exclude_positions=None):
        """``n`` independent ``generate_single`` chains on ONE device batch (the reference's
        `pgen_msa_revised.py:107-115` calls generate_single `seqs_per_template` times, each a batch-1 forward per step).

        Equivalent to ``[self.generate_single(...) for _ in range(n)]``: the position shuffles (Python ``random``) and,
        in replay mode, the Exp(1) variates (torch) are drawn call by call in that order before anything runs -- neither
        stream depends on the model -- and chain c then follows call c's schedule.  Bin sizes depend only on the number
        of positions, so all chains share every step's P."""
        engine = self.model.model.require_engine()
        excluded = {i + 1 for i in (exclude_positions or [])}
        sequence_length = len(seed_msa[0])
        R = len(seed_msa)
        # ---- pre-draw, call-major: schedule[c][pass] = list of bins; noise[c][pass][group] (replay mode)
        schedules, noises = [], []
        for _ in range(n):
            positions = [x for x in range(1, sequence_length + 1) if x not in excluded]
            per_pass, per_pass_noise = [], []
            for pass_num in range(passes):
                random.shuffle(positions)
                bins = partition(positions, steps)
                per_pass.append([list(b) for b in bins])
                if self.rng == "replay":
                    burnin = float("inf") if pass_num < burn_in else 0
                    group_noise, i = [], 0
                    while i < len(bins):
                        j = i
                        while j < len(bins) and len(bins[j]) == len(bins[i]):
                            j += 1
                        group_noise.append(draw_replay_noise(j - i, len(bins[i]), len(self.valid_aa_idx), k, burnin))
                        i = j
                    per_pass_noise.append(group_noise)
            schedules.append(per_pass)
            noises.append(per_pass_noise)
        engine.set_tokens(self.get_init_msa(seed_msa, sequence_length, n))
        for pass_num in range(passes):
            bins0 = schedules[0][pass_num]
            burnin = float("inf") if pass_num < burn_in else 0
            i, g = 0, 0
            while i < len(bins0):
                j = i
                while j < len(bins0) and len(bins0[j]) == len(bins0[i]):
                    j += 1
                P = len(bins0[i])
                pos = np.asarray([[schedules[c][pass_num][st] for c in range(n)] for st in range(i, j)], dtype=np.int32)
                engine.set_schedule(pos, j - i, P, n * P, P, False)
                if self.rng == "replay":
                    stride = noises[0][pass_num][g][1]
                    engine.set_noise(torch.cat([noises[c][pass_num][g][0] for c in range(n)], dim=1), stride)
                else:
                    engine.set_noise(None)
                    engine.set_device_rng(int(torch.randint(0, 2 ** 62, (1,)).item()))
                engine.run_single(0, j - i, burnin, k, None, -1, target_index, self.valid_aa_idx)
                i, g = j, g + 1
        rows = self.untokenize_batch(engine.get_tokens())
        return [rows[c * R + (target_index % R)] for c in range(n)]

    # ------------------------------------------------------------------ scoring (shares the forward)
    def log_likelihood(self, msa, target_index=0, with_masking=True, verbose=False, count_gaps=False,
                       mask_distance=float("inf")) -> Tuple[float, List[float]]:
        return next(self.log_likelihood_batch([msa], target_index, with_masking, verbose, count_gaps, mask_distance))

    def log_likelihood_batch(self, msa_list, target_index=0, with_masking=True, verbose=False, count_gaps=False,
                             mask_distance=float("inf"), batch_size=1) -> Iterator[Tuple[float, List[float]]]:
        """Pseudo-log-likelihood of row ``target_index`` of each MSA (reference esm_msa_sampler.py:319-432): strided
        masking (copy i of the MSA masks positions i, i+n, ... of the target row, n = min(mask_distance, L)), gap
        positions of the target skipped unless ``count_gaps``, values listed copy by copy as the reference does, mean
        accumulated in float32 like its 0-d tensor sum.  Each MSA is scored on its own, so every forward is over
        equal-length rows and no <pad> reaches the engine."""
        alphabet = self.model.alphabet
        gap_tokens = {alphabet.get_idx(c) for c in ESM_MSA_GAP_CHARACTERS}
        start = 1 if alphabet.prepend_bos else 0
        for msa in msa_list:
            cleaned = [self.clean_seed_seq(s) for s in msa]
            toks = self.model.batch_converter([[(str(i), s) for i, s in enumerate(cleaned)]])[2]  # [1, R, C]
            L = len(cleaned[target_index])
            true_toks = toks[0, target_index].tolist()
            values = []

            def collect(lp_row, positions):
                for pos in positions:
                    tok = true_toks[start + pos]
                    if count_gaps or tok not in gap_tokens:
                        values.append(lp_row[start + pos, tok].item())

            n_copies = int(min(mask_distance, L)) if with_masking else 1
            if hasattr(self.model.model, "require_engine"):
                # device path: strided <mask> copies built on the GPU, LM head on the masked rows of the target row
                # only, log_softmax + gather fused into it (Engine.score)
                per_pos = score_strided(self.model.model.require_engine(), toks, true_toks, start, L, n_copies,
                                        batch_size or n_copies, with_masking, row=target_index % toks.shape[1])
                for i in range(n_copies):
                    for pos in range(i, L, n_copies):
                        if count_gaps or true_toks[start + pos] not in gap_tokens:
                            values.append(per_pos[pos])
            elif with_masking:
                bs = batch_size or n_copies
                for b0 in range(0, n_copies, bs):
                    chunk = range(b0, min(b0 + bs, n_copies))
                    t = toks.repeat(len(chunk), 1, 1)
                    for j, i in enumerate(chunk):
                        t[j, target_index, start + i:start + L:n_copies] = alphabet.mask_idx
                    if verbose:
                        print(t[:, target_index])
                    lp = torch.log_softmax(self.model.model(t)["logits"], dim=-1)
                    for j, i in enumerate(chunk):
                        collect(lp[j, target_index], range(i, L, n_copies))
            else:
                lp = torch.log_softmax(self.model.model(toks)["logits"], dim=-1)
                collect(lp[0, target_index], range(L))
            total = np.float32(0.0)
            for v in values:
                total = np.float32(total + np.float32(v))
            yield (float(total / np.float32(len(values))) if values else float("nan"), values)
