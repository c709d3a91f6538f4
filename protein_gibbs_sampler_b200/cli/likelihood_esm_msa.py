"""likelihood_esm_msa.py drop-in (`/root/reference/src/pgen/likelihood_esm_msa.py`): pseudo-log-likelihood of each
query sequence placed on top of a reference alignment and scored by the MSA Transformer.

Per query the context alignment is: a fixed or re-drawn subset of the reference MSA (`random` / `in_order`), or a
mafft alignment of the query's best phmmer hits (`top_hits`); columns where the query has a gap are removed; the top
row is scored with strided masking on the engine (`ESM_MSA_sampler.log_likelihood_batch` -> `Engine.score`).
muscle / phmmer / mafft stay external host programs."""
import argparse
import os
import subprocess
import sys
import tempfile
import textwrap

from tqdm import tqdm

from .. import models
from ..esm_msa_sampler import ESM_MSA_sampler
from ..fasta import (RawAndDefaultsFormatter, SequenceSubsetter, parse_fasta, parse_fasta_string,
                     write_sequential_fasta)
from . import add_weight_flags, build_model
from .likelihood_esm import write_scores
from .pgen_msa_revised import ReferenceDb, delete_msa_cols

model_map = {"esm_msa1": models.ESM_MSA1}


def add_to_msa(msa, new_seq):
    """`muscle -profile`: align `new_seq` to the alignment `msa`; the new sequence comes back first
    (`/root/reference/src/pgen/utils.py:171-207`)."""
    with tempfile.TemporaryDirectory() as tmp:
        msa_path, seq_path = os.path.join(tmp, "msa.fasta"), os.path.join(tmp, "query.fasta")
        write_sequential_fasta(msa_path, msa)
        with open(seq_path, "w") as fh:
            print(f">new_seq\n{new_seq}", file=fh)
        out = subprocess.run(["muscle", "-profile", "-in1", msa_path, "-in2", seq_path], stdout=subprocess.PIPE,
                             stderr=subprocess.PIPE, encoding="utf-8")
    names, seqs = parse_fasta_string(out.stdout, True)
    if "new_seq" not in names:
        print(names, out.stdout, out.stderr, sep="\n", file=sys.stderr)
        raise ValueError("muscle output does not contain the added sequence")
    i = names.index("new_seq")
    return [seqs[i]] + seqs[:i] + seqs[i + 1:]


class ContextBuilder:
    """Query name -> alignment with the query on top (what the reference's `get_in_msa` closure returns, :77-107)."""

    def __init__(self, in_seqs, in_msas=None, reference_msa=None, subset_strategy="random", alignment_size=sys.maxsize,
                 subset_random_seed=None, redraw=False, unaligned_queries=False, keep_identical=False):
        self.in_seqs, self.in_msas = in_seqs, in_msas
        self.reference_msa = reference_msa
        self.strategy, self.size, self.seed = subset_strategy, alignment_size, subset_random_seed
        self.redraw, self.unaligned_queries, self.keep_identical = redraw, unaligned_queries, keep_identical
        self.fixed = self.db = None
        if in_msas:
            return
        if subset_strategy == "top_hits":
            self.db = ReferenceDb(reference_msa)
        else:
            self.fixed = self.draw()

    def close(self):
        if self.db is not None:
            self.db.__exit__()
            self.db = None

    def draw(self):
        return SequenceSubsetter.subset(self.reference_msa, self.size, strategy=self.strategy, random_seed=self.seed)

    def __call__(self, name):
        if self.in_msas:
            return self.in_msas[name]
        seq = self.in_seqs[name]
        if self.db is not None:
            return self.db.alignment_with_top_hits(name, seq, self.size, self.keep_identical)
        context = self.fixed
        if self.redraw:
            context = self.draw()
            if self.seed is not None:
                self.seed += 1000000
        return add_to_msa(context, seq) if self.unaligned_queries else [seq] + list(context)


def main(input_h, output_h, masking_off, sampler, reference_msa_handle=None, in_msas=None, delete_insertions=False,
         batch_size=1, subset_strategy="random", alignment_size=sys.maxsize, subset_random_seed=None, redraw=False,
         unaligned_queries=False, mask_distance=float("inf"), csv=False, positionwise=None, keep_identical=False,
         show_progress_bar=True):
    """`in_msas`: {name: alignment with the query on top}; when given, `reference_msa_handle` is ignored."""
    clean_flag = "delete" if delete_insertions else "upper"
    sep = "," if csv else "\t"
    reference_msa = None if in_msas else parse_fasta(reference_msa_handle, clean=clean_flag)
    in_seqs = dict(zip(*parse_fasta(input_h, return_names=True, clean=clean_flag)))
    context_of = ContextBuilder(in_seqs, in_msas, reference_msa, subset_strategy, alignment_size, subset_random_seed,
                                redraw, unaligned_queries, keep_identical)
    positionwise_h = open(positionwise, "w") if positionwise is not None else None
    try:
        print(f"id{sep}esm-msa", file=output_h)
        if positionwise_h is not None:
            print(f"id{sep}esm-msa", file=positionwise_h)
        names = list(in_seqs)
        for i in tqdm(range(0, len(names), batch_size), disable=not show_progress_bar):
            msas = []
            for name in names[i:i + batch_size]:
                msa = context_of(name)
                query_gaps = [c for c, ch in enumerate(msa[0]) if ch == "-"]   # columns the query does not occupy
                msas.append(delete_msa_cols(msa, query_gaps))
            scores = sampler.log_likelihood_batch(msas, with_masking=not masking_off, count_gaps=False,
                                                  mask_distance=mask_distance, batch_size=batch_size)
            write_scores(names[i:i + batch_size], scores, output_h, positionwise_h, sep)
    finally:
        if positionwise_h is not None:
            positionwise_h.close()
        context_of.close()


def build_parser():
    parser = argparse.ArgumentParser(
        description=textwrap.dedent("""Pseudo-log-likelihood of every query sequence under the MSA Transformer (B200
            engine), each query scored as the top row of a context alignment built from --reference_msa.

            Output: one line per query, name and score, tab (or comma) separated.
            """), formatter_class=RawAndDefaultsFormatter)
    parser.add_argument("-o", type=str, default=None, help="output table (default: stdout)")
    parser.add_argument("-i", default=None, help="fasta of the queries; default stdin")
    parser.add_argument("--reference_msa", default=None, required=True,
                        help="context alignment (fasta / a2m); with --subset_strategy top_hits an UNALIGNED fasta that is "
                             "searched per query")
    parser.add_argument("--device", type=str, default="gpu", help="gpu (cuda:0) or cuda:[int]; the engine has no cpu path")
    parser.add_argument("--masking_off", action="store_true", default=False, help="score every position from ONE unmasked forward")
    parser.add_argument("--delete_insertions", action="store_true", default=False,
                        help="drop a2m insertion columns (lower case, '.') instead of upper-casing them / turning '.' into '-'")
    parser.add_argument("--alignment_size", type=int, default=sys.maxsize,
                        help="context rows taken from the reference (31-255 is sensible); default: all of them")
    parser.add_argument("--keep_identical", action="store_true", default=False,
                        help="top_hits: keep hits identical to the query")
    parser.add_argument("--batch_size", type=int, default=1, help="queries per call and masked MSA copies per forward")
    parser.add_argument("--subset_strategy", default="random", choices=["in_order", "random", "top_hits"],
                        help="random: a shuffle's first rows; in_order: the first rows; top_hits: phmmer the query against "
                             "the references and mafft-align it with its best hits")
    parser.add_argument("--subset_random_seed", default=None, type=int,
                        help="seed of the random subset (+1000000 after every draw)")
    parser.add_argument("--redraw", action="store_true", default=False,
                        help="random: draw a fresh subset for every query instead of one for all")
    parser.add_argument("--unaligned_queries", action="store_true", default=False,
                        help="queries are not in the reference's column frame: add each one with muscle -profile")
    parser.add_argument("--mask_distance", type=int, default=None,
                        help="mask every mask_distance-th position of a copy at once; default: one position per copy")
    parser.add_argument("--csv", action="store_true", default=False, help="comma instead of tab separated")
    parser.add_argument("--positionwise", type=str, default=None,
                        help="also write per-position values here: name, then the ';'-joined list rounded to 3 decimals")
    add_weight_flags(parser)
    return parser


def cli(argv=None):
    args = build_parser().parse_args(argv)
    args.model = "esm_msa1"   # the only MSA model, as in the reference (:197)
    if args.redraw and args.subset_strategy == "in_order":
        raise ValueError("redraw is set, but subset_strategy is 'in_order', so all the draws will be the same. "
                         "That's probably not what you're trying to do.")
    mask_distance = float("inf") if args.mask_distance is None else args.mask_distance
    if mask_distance < 1:
        raise ValueError("mask distance must be an integer >= 1.")
    input_handle = open(args.i, "r") if args.i is not None else sys.stdin
    output_handle = open(args.o, "w") if args.o is not None else sys.stdout
    try:
        sampler = ESM_MSA_sampler(build_model(model_map, args), device=args.device)
        with open(args.reference_msa, "r") as reference_msa_handle:
            main(input_handle, output_handle, args.masking_off, sampler, reference_msa_handle,
                 delete_insertions=args.delete_insertions, batch_size=args.batch_size,
                 subset_strategy=args.subset_strategy, alignment_size=args.alignment_size,
                 subset_random_seed=args.subset_random_seed, redraw=args.redraw,
                 unaligned_queries=args.unaligned_queries, mask_distance=mask_distance, csv=args.csv,
                 positionwise=args.positionwise, keep_identical=args.keep_identical)
    finally:
        if args.i is not None:
            input_handle.close()
        if args.o is not None:
            output_handle.close()


if __name__ == "__main__":
    cli()
