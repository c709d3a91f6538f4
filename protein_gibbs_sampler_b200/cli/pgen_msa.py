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
from . import add_weight_flags, build_model, spec_args, spec_lines
from .pgen_esm import EPILOG

model_map = {"esm_msa1": models.ESM_MSA1}


def main(input_h, output_p, args):
    clean_flag = "delete" if args.delete_insertions else "upper"
    gibbs_sampler = ESM_MSA_sampler(build_model(model_map, args), device=args.device)
    with open(output_p / "specification.tsv", "w") as echo_h:
        for name, arg_text, msa_path in spec_lines(input_h, echo_h, 3, "name, line_args, input_msa"):
            line_args = spec_args(arg_text)
            input_msa = parse_fasta(msa_path, clean=clean_flag)
            rows = len(input_msa) if args.alignment_size == sys.maxsize else args.alignment_size
            sequences = []
            # every round draws a fresh subset of the seed alignment and resamples ALL of its rows
            for _ in trange(math.ceil(args.num_output_sequences / rows)):
                seed_msa = SequenceSubsetter.subset(input_msa, rows, args.keep_first_sequence, args.subset_strategy)
                sequences += gibbs_sampler.generate(n_samples=len(seed_msa), seed_msa=seed_msa,
                                                    batch_size=args.batch_size, show_progress_bar=False, **line_args)
            write_sequential_fasta(output_p / (name + ".fasta"), sequences[:args.num_output_sequences])


def build_parser():
    parser = argparse.ArgumentParser(
        description=textwrap.dedent("""Gibbs-sample new rows of a protein alignment from the MSA Transformer (B200 engine).

            One run per input line:  <run name> TAB <python dict of sampler arguments> TAB <seed alignment>
            """),
        epilog=EPILOG, formatter_class=RawAndDefaultsFormatter)
    parser.add_argument("-o", default=".", help="output directory (created if missing)")
    parser.add_argument("-i", default=None, help="specification file, one run per line: "
                        "name TAB dict of sampler arguments TAB seed alignment (fasta / a2m); default stdin")
    parser.add_argument("--batch_size", type=int, default=1,
                        help="MSAs resampled together on the GPU (the reference script only allows 1, pgen_msa.py:78)")
    parser.add_argument("--num_output_sequences", type=int, default=1, help="sequences written per run")
    parser.add_argument("--device", type=str, default="gpu", help="gpu (cuda:0) or cuda:[int]; the engine has no cpu path")
    parser.add_argument("--model", type=str, default="esm_msa1", choices=sorted(model_map), help="model triple (architecture + alphabet)")
    parser.add_argument("--delete_insertions", action="store_true", default=False,
                        help="drop a2m insertion columns (lower case, '.') instead of upper-casing them / turning '.' into '-'")
    parser.add_argument("--alignment_size", type=int, default=sys.maxsize,
                        help="rows of the seed alignment given to the model (32-256 is sensible); default: all of them")
    parser.add_argument("--keep_first_sequence", action="store_true", default=False,
                        help="always keep row 0 and choose only the other rows with --subset_strategy")
    parser.add_argument("--subset_strategy", default="random", choices=sorted(SequenceSubsetter.subset_strategies),
                        help="how the --alignment_size rows are chosen")
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
