"""pgen_esm_from_fasta.py drop-in (`/root/reference/src/pgen/pgen_esm_from_fasta.py`): TSV of
`name <tab> dict of sampler arguments <tab> fasta of seed sequences`; every output sequence starts from a seed drawn
with `random.choice` and un-aligned.

The reference runs one `generate(n_samples=1)` per output sequence (:27-33).  Here the draws are made in the same
order but the chains run together, one device batch per seed length (`ESM_sampler.generate_many`)."""
import argparse
import random
import sys
import textwrap
from pathlib import Path

from .. import models
from ..esm_sampler import ESM_sampler
from ..fasta import RawAndDefaultsFormatter, add_gaps_back, parse_fasta, unalign, write_sequential_fasta
from . import add_weight_flags, build_model, spec_args
from .pgen_esm import EPILOG

model_map = {"esm1b": models.ESM1b, "esm6": models.ESM6, "esm12": models.ESM12, "esm34": models.ESM34,
             "esm2_t6_8M": models.ESM2_t6_8M, "esm2_t30_150M": models.ESM2_t30_150M,
             "esm2_t33_650M": models.ESM2_t33_650M}


def sample_from_seeds(sampler, seeds, n_outputs, line_args, keep_gap_positions=False):
    """n_outputs sequences, each grown from `random.choice(seeds)` with its gaps removed (and put back afterwards if
    keep_gap_positions)."""
    gap_masks = []

    def draw():
        for _ in range(n_outputs):
            seed, gap_mask = unalign(random.choice(seeds))
            gap_masks.append(gap_mask)
            yield seed

    sequences = sampler.generate_many(draw(), **line_args)
    if keep_gap_positions:
        sequences = [add_gaps_back(s, m) for s, m in zip(sequences, gap_masks)]
    return sequences


def main(input_h, output_p, args, sampler=None):
    if sampler is None:
        sampler = ESM_sampler(build_model(model_map, args), device=args.device)
    with open(output_p / "specification.tsv", "w") as output_h:
        for line in input_h:
            line = line.strip()
            if not line:
                continue
            fields = line.split("\t")
            if len(fields) != 3:   # the reference skips such lines silently (:22)
                continue
            print("\t".join(fields))
            print("\t".join(fields), file=output_h)
            seeds = parse_fasta(fields[2], clean=None)
            sequences = sample_from_seeds(sampler, seeds, args.num_output_sequences, spec_args(fields[1]),
                                          args.keep_gap_positions)
            write_sequential_fasta(output_p / (fields[0] + ".fasta"), sequences)


def build_parser():
    parser = argparse.ArgumentParser(
        description=textwrap.dedent("""Samples from an ESM BERT model to generate new protein sequences.

            Input should be a tab separated file where columns are:
            sample name, dict of sampler arguments, fasta of seed sequences
            """),
        epilog=EPILOG.replace("seed_seq: protein sequence (or list of sequences) to start from\n", ""),
        formatter_class=RawAndDefaultsFormatter)
    parser.add_argument("-o", default=".", help="a directory to save the outputs to.")
    parser.add_argument("-i", default=None, help="tab separated file where the columns are as follows: [sample name] "
                                                 "\\t [dict of arguments for the sampler] \\t [path to fasta file].")
    parser.add_argument("--batch_size", type=int, default=1, choices=[1],
                        help="kept for compatibility (must be 1): chains of equal length are batched on the device anyway.")
    parser.add_argument("--num_output_sequences", type=int, default=1, help="total number of sequences to generate.")
    parser.add_argument("--device", type=str, default="gpu", help="gpu (cuda:0) or cuda:[int]; the engine has no cpu path")
    parser.add_argument("--model", type=str, default="esm1b", choices=sorted(model_map), help="which model to use")
    parser.add_argument("--keep_gap_positions", action="store_true", default=False,
                        help="If set, remember where the gaps are in the seed and add them back into the same "
                             "positions of the generated sequence.")
    add_weight_flags(parser)
    return parser


def cli(argv=None):
    args = build_parser().parse_args(argv)
    input_handle = open(args.i, "r") if args.i is not None else sys.stdin
    output_path = Path(args.o)
    output_path.mkdir(exist_ok=True)
    try:
        main(input_handle, output_path, args)
    finally:
        if args.i is not None:
            input_handle.close()


if __name__ == "__main__":
    cli()
