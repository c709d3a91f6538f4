"""CPU restatement of the reference's Gibbs loop (TEST INFRASTRUCTURE / CPU BASELINE, see oracle/__init__.py).

Restates, statement for statement but in this project's own words, what
/root/reference/src/pgen/esm_sampler.py does on its hot path:
  generate_step                 :8-45     -> ``generate_step``
  ESM_sampler.generate          :173-240  -> ``esm_generate``      (loop body :209-234)
  ESM_MSA_sampler.generate      :185-253  -> ``msa_generate``      (loop body :221-248)
  ESM_MSA_sampler.generate_single :101-147 -> ``msa_generate_single``
It keeps the reference's execution shape -- one eager forward per iteration, then one ``generate_step`` per
residue with a scalar write-back -- because it doubles as the timed CPU baseline of ``bench.py``.
Pinned against the reference itself (imported from /root/reference/src) by tests/golden/make_golden.py:
identical outputs under identical Python/torch seeds.
"""
import math
import random

import torch

ESM_AA = "ACDEFGHIKLMNPQRSTVWY"
MSA_AA = "-ACDEFGHIKLMNPQRSTVWY"


def generate_step(out, gen_idx, temperature=None, top_k=0, sample=False, valid_idx=None):
    logits = out[gen_idx]
    if temperature is not None:
        logits = logits / temperature
    if valid_idx is None:
        valid_idx = list(range(len(logits)))
    candidates = logits[valid_idx]
    if sample or top_k <= 0 or top_k > len(candidates):
        top_k = len(candidates)
    vals, order = candidates.topk(top_k)
    pick = torch.distributions.categorical.Categorical(logits=vals).sample()
    return torch.tensor(valid_idx[order[pick]])


def _clean(seq, allowed):
    seq = seq.upper()
    bad = set(seq) - set(allowed)
    if bad:
        raise Exception("Invalid input character: " + ",".join(bad))
    return seq


def _indexes(indexes, leader_length, max_len, rollover, as_list):
    if indexes is not None:
        return indexes, -1
    idx = range(1, max_len + 1)
    if as_list:
        idx = list(idx)
    if rollover:
        return idx, -1
    return idx[leader_length:], leader_length - 1


def _in_order(indexes, cursor, n):
    out = []
    for _ in range(n):
        cursor = (cursor + 1) % len(indexes)
        out.append(indexes[cursor])
    return cursor, out


def esm_generate(model, n_samples, seed_seq, batch_size=1, in_order=False, max_len=None, leader_length=0,
                 leader_length_percent=None, top_k=0, temperature=None, num_iters=10, burnin=float("inf"),
                 mask=True, num_positions=0, num_positions_percent=None, indexes=None, rollover_from_start=False,
                 on_iteration=None):
    """``model`` is the duck-typed triple (model / alphabet / batch_converter)."""
    a = model.alphabet
    valid = sorted(a.get_idx(t) for t in ESM_AA)
    if isinstance(seed_seq, str):
        seq_len = len(seed_seq)
    elif isinstance(seed_seq, list):
        seq_len = max(len(s) for s in seed_seq)
    else:
        raise ValueError("Unknown seed sequence format, expecting str or list")
    if max_len is None:
        max_len = seq_len
    if num_positions_percent is not None:
        num_positions = int(max_len * (num_positions_percent / 100))
    num_positions = max(num_positions, 0)
    if leader_length_percent is not None:
        leader_length = int(max_len * (leader_length_percent / 100))
    leader_length = max(leader_length, 0)
    out_seqs = []
    n_batches = math.ceil(n_samples / batch_size)
    with torch.no_grad():
        for bn in range(n_batches):
            if isinstance(seed_seq, list):
                chosen = random.choices(seed_seq, k=batch_size)
                rows = [(str(i), _clean(s, ESM_AA) + "<mask>" * (max_len - len(s))) for i, s in enumerate(chosen)]
            else:
                fill = "<mask>" * (max_len - len(seed_seq))
                rows = [(str(i), _clean(seed_seq, ESM_AA) + fill) for i in range(batch_size)]
            batch = model.batch_converter(rows)[2]
            indexes, cursor = _indexes(indexes, leader_length, max_len, rollover_from_start, as_list=False)
            num_positions = min(num_positions, len(indexes))
            for ii in range(num_iters):
                if num_positions > 0:
                    if in_order:
                        cursor, picked = _in_order(indexes, cursor, num_positions)
                        targets = [picked] * batch_size
                    else:
                        targets = [random.sample(indexes, num_positions) for _ in range(batch_size)]
                else:
                    targets = [indexes] * batch_size
                if mask:
                    for b in range(batch_size):
                        for kk in targets[b]:
                            batch[b][kk] = a.mask_idx
                logits = model.model(batch)["logits"]
                for b in range(batch_size):
                    for kk in targets[b]:
                        batch[b][kk] = generate_step(logits[b], kk, temperature=temperature, top_k=top_k,
                                                     sample=(ii < burnin), valid_idx=valid)
                if on_iteration is not None:
                    on_iteration(ii, batch)
            lo = 1 if a.prepend_bos else 0
            hi = -1 if a.append_eos else None
            strs = ["".join(a.get_tok(t) for t in row.tolist()[lo:hi]) for row in batch]
            out_seqs += strs[0:n_samples - len(out_seqs)] if bn == n_batches - 1 else strs
    return out_seqs


def msa_generate(model, n_samples, seed_msa, batch_size=1, in_order=False, max_len=None, leader_length=0,
                 leader_length_percent=None, top_k=0, temperature=None, num_iters=10, burnin=float("inf"),
                 mask=True, num_positions=0, num_positions_percent=None, indexes=None, rollover_from_start=False):
    a = model.alphabet
    valid = sorted(a.get_idx(t) for t in MSA_AA)
    R, seq_len = len(seed_msa), len(seed_msa[0])
    rounds = math.ceil(n_samples / R / batch_size)
    if num_positions_percent is not None:
        num_positions = int(seq_len * (num_positions_percent / 100))
    num_positions = max(num_positions, 0)
    if leader_length_percent is not None:
        leader_length = int(seq_len * (leader_length_percent / 100))
    leader_length = max(leader_length, 0)
    if max_len is None:
        max_len = seq_len
    out_seqs = []
    with torch.no_grad():
        for rnd in range(rounds):
            padded = [(str(i), _clean(s, MSA_AA) + "<mask>" * (max_len - len(s))) for i, s in enumerate(seed_msa)]
            batch = model.batch_converter([padded] * batch_size)[2]
            indexes, cursor = _indexes(indexes, leader_length, max_len, rollover_from_start, as_list=True)
            num_positions = min(num_positions, len(indexes))
            for ii in range(num_iters):
                if num_positions > 0:
                    if in_order:
                        cursor, picked = _in_order(indexes, cursor, num_positions)
                        targets = [[picked] * R for _ in range(batch_size)]
                    else:
                        targets = [[random.sample(indexes, num_positions) for _ in range(R)]
                                   for _ in range(batch_size)]
                else:
                    targets = [[indexes] * R for _ in range(batch_size)]
                if mask:
                    for b in range(batch_size):
                        for r in range(R):
                            for kk in targets[b][r]:
                                batch[b][r][kk] = a.mask_idx
                logits = model.model(batch)["logits"]
                for b in range(batch_size):
                    for r in range(R):
                        for kk in targets[b][r]:
                            batch[b][r][kk] = generate_step(logits[b][r], kk, temperature=temperature, top_k=top_k,
                                                            sample=(ii < burnin), valid_idx=valid)
            strs = ["".join(a.get_tok(t) for t in row.tolist()[1:]) for msa in batch for row in msa]
            out_seqs += strs[0:n_samples - len(out_seqs)] if rnd == rounds - 1 else strs
    return out_seqs


def partition(items, bins):
    bins = min(bins, len(items))
    base, extra = divmod(len(items), bins) if bins else (0, 0)
    out, lo = [], 0
    for i in range(bins):
        hi = lo + base + (1 if i < extra else 0)
        out.append(list(items[lo:hi]))
        lo = hi
    return out


def msa_generate_single(model, seed_msa, steps=10, passes=3, burn_in=1, target_index=0, k=1,
                        exclude_positions=None):
    a = model.alphabet
    valid = sorted(a.get_idx(t) for t in MSA_AA)
    excluded = {i + 1 for i in (exclude_positions or [])}
    L = len(seed_msa[0])
    positions = [p for p in range(1, L + 1) if p not in excluded]
    with torch.no_grad():
        padded = [(str(i), _clean(s, MSA_AA) + "<mask>" * (L - len(s))) for i, s in enumerate(seed_msa)]
        batch = model.batch_converter([padded])[2]
        for pass_num in range(passes):
            random.shuffle(positions)
            for group in partition(positions, steps):
                for kk in group:
                    batch[0][-1][kk] = a.mask_idx          # reference masks the LAST row (:133)
                logits = model.model(batch)["logits"]
                for kk in group:
                    batch[0][target_index][kk] = generate_step(logits[0][target_index], kk, top_k=k,
                                                               sample=(pass_num < burn_in), valid_idx=valid)
        rows = ["".join(a.get_tok(t) for t in row.tolist()[1:]) for row in batch[0]]
    return rows[target_index]
