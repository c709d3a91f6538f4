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
from . import add_weight_flags, build_model, spec_args, spec_lines
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
    with open(output_p / "specification.tsv", "w") as echo_h:
        # (the reference skips malformed lines silently here, :22)
        for name, arg_text, seeds_path in spec_lines(input_h, echo_h, 3, "name, line_args, seed fasta", complain=False):
            sequences = sample_from_seeds(sampler, parse_fasta(seeds_path, clean=None), args.num_output_sequences,
                                          spec_args(arg_text), args.keep_gap_positions)
            write_sequential_fasta(output_p / (name + ".fasta"), sequences)


def build_parser():
    parser = argparse.ArgumentParser(
        description=textwrap.dedent("""Gibbs-sample new protein sequences, each started from a randomly chosen sequence of
            a seed fasta (B200 engine).

            One run per input line:  <run name> TAB <python dict of sampler arguments> TAB <seed fasta>
            """),
        epilog=EPILOG.replace("seed_seq: protein sequence (or list of sequences) to start from\n", ""),
        formatter_class=RawAndDefaultsFormatter)
    parser.add_argument("-o", default=".", help="output directory (created if missing)")
    parser.add_argument("-i", default=None, help="specification file (name TAB dict TAB seed fasta per line); default stdin")
    parser.add_argument("--batch_size", type=int, default=1, choices=[1],
                        help="kept for compatibility (must be 1): chains of equal length are batched on the device anyway.")
    parser.add_argument("--num_output_sequences", type=int, default=1, help="sequences written per run")
    parser.add_argument("--device", type=str, default="gpu", help="gpu (cuda:0) or cuda:[int]; the engine has no cpu path")
    parser.add_argument("--model", type=str, default="esm1b", choices=sorted(model_map), help="model triple (architecture + alphabet)")
    parser.add_argument("--keep_gap_positions", action="store_true", default=False,
                        help="put the seed's gap characters back at their columns in the generated sequence")
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
