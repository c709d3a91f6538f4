"""FASTA / alignment text helpers the command-line drop-ins need.

Host-side string handling only (SURVEY.md section 2, row 10 is out of scope as a subsystem): just the functions
the three CLIs call, with the behaviour the reference's tests pin (`/root/reference/test/test_utils.py:26-110`):
`parse_fasta` clean modes (`/root/reference/src/pgen/utils.py:87-168`), `write_sequential_fasta` (`:221-229`),
`SequenceSubsetter.subset` (`:316-356`), `unalign` / `add_gaps_back` (`:42-85`).
"""
import argparse
import io
import random
import string


class RawAndDefaultsFormatter(argparse.ArgumentDefaultsHelpFormatter, argparse.RawDescriptionHelpFormatter):
    """argparse formatter used by every CLI: raw description text plus default values."""


def _handle(path_or_handle, mode="r"):
    """(file object, opened_here).  Anything `open` rejects with TypeError is taken to be a handle already."""
    try:
        return open(path_or_handle, mode), True
    except TypeError:
        return path_or_handle, False


_CLEAN_TABLES = {
    # a2m insertions (lowercase, '.') and '*' dropped: alignment columns only
    "delete": str.maketrans(dict.fromkeys(string.ascii_lowercase + ".*")),
    # keep the length: upper-case first, then '*' dropped and '.' -> '-'
    "upper": str.maketrans({"*": None, ".": "-"}),
    # plain sequence: upper-case first, then every gap-like character dropped
    "unalign": str.maketrans({"*": None, ".": None, "-": None}),
}


def parse_fasta(filename, return_names=False, clean=None, full_name=False):
    """Sequences (or (names, sequences)) of a FASTA / a2m file or open handle.

    clean: None | 'upper' | 'delete' | 'unalign'; names stop at the first whitespace unless full_name."""
    if clean is not None and clean not in _CLEAN_TABLES:
        raise ValueError(f"unrecognized input for clean parameter: {clean}")
    fh, opened = _handle(filename)
    names, seqs = [], []
    try:
        for raw in fh:
            line = raw.strip()
            if not line:
                continue
            if line.startswith(">"):
                names.append(line[1:] if full_name else line.split(None, 1)[0][1:])
                seqs.append([])
            elif seqs:
                seqs[-1].append(line)
    finally:
        if opened:
            fh.close()
    out = ["".join(parts) for parts in seqs]
    if clean == "delete":
        out = [s.translate(_CLEAN_TABLES["delete"]) for s in out]
    elif clean is not None:
        out = [s.upper().translate(_CLEAN_TABLES[clean]) for s in out]
    return (names, out) if return_names else out


def parse_fasta_string(fasta_string, return_names=False):
    return parse_fasta(io.StringIO(fasta_string), return_names)


def write_sequential_fasta(path, sequences):
    """Records named 0..len-1, one line per sequence."""
    fh, opened = _handle(path, "w")
    for i, seq in enumerate(sequences):
        print(f">{i}\n{seq}", file=fh)
    if opened:
        fh.close()


def write_partitioned_fasta(path, sequences):
    """`sequences`: {category: [seq, ...]} -> records named category_index."""
    with open(path, "w") as fh:
        for category, seqs in sequences.items():
            for i, seq in enumerate(seqs):
                print(f">{category}_{i}\n{seq}", file=fh)


def unalign(sequence):
    """Upper-case letters of `sequence`, and a mask (None where a letter was, else the dropped character)."""
    letters, mask = [], []
    for c in sequence.upper():
        if c in string.ascii_uppercase:
            letters.append(c)
            mask.append(None)
        else:
            mask.append(c)
    return "".join(letters), mask


def add_gaps_back(sequence, gap_mask):
    """Inverse of `unalign`: letters of `sequence` go where the mask holds None."""
    it = iter(sequence)
    return "".join(next(it) if c is None else c for c in gap_mask)


class SequenceSubsetter:
    subset_strategies = {"random", "in_order"}

    @classmethod
    def subset(cls, seq_list, n, keep_first=False, strategy="random", random_seed=None):
        """n members of seq_list: the first n ('in_order') or a shuffle's first n ('random', own Random(seed));
        keep_first pins seq_list[0] and draws the other n-1 from the rest."""
        if n <= 0:
            return []
        if strategy not in cls.subset_strategies:
            raise ValueError(f"sampler strategy {strategy} not recognized, must be one of {cls.subset_strategies}")
        head, pool = ([seq_list[0]], list(seq_list[1:])) if keep_first else ([], list(seq_list))
        if strategy == "random":
            random.Random(random_seed).shuffle(pool)
        return head + pool[:n - len(head)]
