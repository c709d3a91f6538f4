"""pgen_msa.py drop-in (`/root/reference/src/pgen/pgen_msa.py`): TSV of `name <tab> dict of sampler arguments <tab>
seed alignment (fasta / a2m)` -> subsets of the alignment -> `ESM_MSA_sampler.generate` on the GPU -> FASTA."""
import argparse
import math
import sys
import textwrap
from pathlib import Path

from tqdm import trange

from .. import models
from ..esm_msa_sampler import ESM_MSA_sampler
from ..fasta import RawAndDefaultsFormatter, SequenceSubsetter, parse_fasta, write_sequential_fasta
from . import add_weight_flags, build_model, spec_args
from .pgen_esm import EPILOG

model_map = {"esm_msa1": models.ESM_MSA1}


def main(input_h, output_p, args):
    clean_flag = "delete" if args.delete_insertions else "upper"
    gibbs_sampler = ESM_MSA_sampler(build_model(model_map, args), device=args.device)
    with open(output_p / "specification.tsv", "w") as output_h:
        for line in input_h:
            line = line.strip()
            if not line:
                continue
            fields = line.split("\t")
            if len(fields) != 3:
                print(f"Expected 3 values in specification file (name, line_args, input_msa), got {len(fields)}")
                print("\t".join(fields))
                continue
            print("\t".join(fields))
            print("\t".join(fields), file=output_h)
            name, line_args = fields[0], spec_args(fields[1])
            input_msa = parse_fasta(fields[2], clean=clean_flag)
            alignment_size = len(input_msa) if args.alignment_size == sys.maxsize else args.alignment_size
            sequences = []
            for _ in trange(math.ceil(args.num_output_sequences / alignment_size)):
                batch_msa = SequenceSubsetter.subset(input_msa, alignment_size, args.keep_first_sequence,
                                                     args.subset_strategy)
                sequences += gibbs_sampler.generate(n_samples=len(batch_msa), seed_msa=batch_msa,
                                                    batch_size=args.batch_size, show_progress_bar=False, **line_args)
            write_sequential_fasta(output_p / (name + ".fasta"), sequences[0:args.num_output_sequences])


def build_parser():
    parser = argparse.ArgumentParser(
        description=textwrap.dedent("""Samples from the ESM-MSA model to generate new protein sequences.

            Input should be a tab separated file where columns are:
            sample name, dict of sampler arguments, fasta of seed sequences
            """),
        epilog=EPILOG, formatter_class=RawAndDefaultsFormatter)
    parser.add_argument("-o", default=".", help="a directory to save the outputs in")
    parser.add_argument("-i", default=None, help="tab separated file where the columns are as follows: [sample name] "
                        "\\t [dict of arguments for the sampler] \\t [path to seed msa in fasta or a2m format].")
    parser.add_argument("--batch_size", type=int, default=1,
                        help="batch size for sampling (msa instances per iteration).  The reference restricts this to 1 "
                             "(pgen_msa.py:78); the engine batches whole MSAs, so any value is accepted.")
    parser.add_argument("--num_output_sequences", type=int, default=1, help="total number of sequences to generate.")
    parser.add_argument("--device", type=str, default="gpu", help="gpu (cuda:0) or cuda:[int]; the engine has no cpu path")
    parser.add_argument("--model", type=str, default="esm_msa1", choices=sorted(model_map), help="which model to use")
    parser.add_argument("--delete_insertions", action="store_true", default=False,
                        help="If set, then remove all lowercase and '.' characters from input sequences. Default: "
                             "convert lower to upper and '.' to '-'.")
    parser.add_argument("--alignment_size", type=int, default=sys.maxsize,
                        help="Sample this many sequences from the input alignment before doing gibbs sampling, "
                             "recommended values are 32-256. Default: the entire input alignment.")
    parser.add_argument("--keep_first_sequence", action="store_true", default=False,
                        help="If set, then keep the first sequence and sample the rest according to subset_strategy.")
    parser.add_argument("--subset_strategy", default="random", choices=sorted(SequenceSubsetter.subset_strategies),
                        help="How to subset the input alignment to get it to the desired size.")
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
