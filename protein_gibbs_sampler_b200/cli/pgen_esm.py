"""pgen_esm.py drop-in (`/root/reference/src/pgen/pgen_esm.py`): TSV of `name <tab> dict of sampler arguments`
-> `ESM_sampler.generate` on the GPU -> one FASTA per line plus an echo of the specification."""
import argparse
import sys
import textwrap
from pathlib import Path

from .. import models
from ..esm_sampler import ESM_sampler
from ..fasta import RawAndDefaultsFormatter, write_sequential_fasta
from . import add_weight_flags, build_model, spec_args

# the reference's names (pgen_esm.py:10) plus the ESM-2 family (BASELINE configs 1 and 4)
model_map = {"esm1b": models.ESM1b, "esm6": models.ESM6, "esm12": models.ESM12, "esm34": models.ESM34,
             "esm1v": models.ESM1v, "esm2_t6_8M": models.ESM2_t6_8M,
             "esm2_t30_150M": models.ESM2_t30_150M, "esm2_t33_650M": models.ESM2_t33_650M}

EPILOG = """
Available sampler arguments:

seed_seq: protein sequence (or list of sequences) to start from
in_order: if True then cycle through the positions in order, otherwise randomly select positions each iteration.
max_len: maximum size of each generated sequence. If None, then use the length of the longest input sequence.
leader_length: don't overwrite this many amino acids at the beginning of the sequence.
leader_length_percent: if not None, then will set leader_length = int(len(seed_seq)*(leader_length_percent / 100))
top_k: if >0, only sample from the top k most probable AAs
temperature: higher numbers will mean there is a lower penalty for low-scoring amino acids.
num_iters: how many times to run the forward loop for every batch.
burnin: during burn-in period, sample from full distribution; afterwards sample from top_k, set to 0 to never sample
        from full distribution (always take from top_k), or inf to always sample from full distribution.
num_positions: generate new AAs for this many positions each iteration. If 0, then generate for all target positions.
num_positions_percent: If not None, then set num_positions = int(len(seed_seq)*(num_positions_percent / 100))
indexes: positions of the input sequence to modify. 1-indexed, if None then all positions after the leader.
"""


def main(input_h, output_p, args):
    sampler = ESM_sampler(build_model(model_map, args), device=args.device)
    with open(output_p / "specification.tsv", "w") as output_h:
        for line in input_h:
            line = line.strip()
            if not line:
                continue
            fields = line.split("\t")
            if len(fields) != 2:
                print(f"Expected 2 values in specification file (name, line_args), got {len(fields)}")
                print("\t".join(fields))
                continue
            print("\t".join(fields))
            print("\t".join(fields), file=output_h)
            name, line_args = fields[0], spec_args(fields[1])
            sequences = sampler.generate(args.num_output_sequences, batch_size=args.batch_size, **line_args)
            write_sequential_fasta(output_p / (name + ".fasta"), sequences)


def build_parser():
    parser = argparse.ArgumentParser(
        description=textwrap.dedent("""Samples from an ESM BERT model to generate new protein sequences.

            Input should be a tab separated file where columns are:
            sample name, dict of sampler arguments
            """),
        epilog=EPILOG, formatter_class=RawAndDefaultsFormatter)
    parser.add_argument("-o", default=".", help="a directory to save the outputs to.")
    parser.add_argument("-i", default=None, help="tab separated file where the columns are as follows: "
                                                 "[sample name] \\t [dict of arguments for the sampler].")
    parser.add_argument("--batch_size", type=int, default=1, help="batch size for sampling (sequences per iteration).")
    parser.add_argument("--num_output_sequences", type=int, default=1, help="total number of sequences to generate.")
    parser.add_argument("--device", type=str, default="gpu", help="gpu (cuda:0) or cuda:[int]; the engine has no cpu path")
    parser.add_argument("--model", type=str, default="esm1b", choices=sorted(model_map), help="which model to use")
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
