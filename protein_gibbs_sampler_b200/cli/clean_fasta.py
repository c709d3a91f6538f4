"""clean_fasta.py drop-in (`/root/reference/src/pgen/clean_fasta.py`): re-write a FASTA through one of
`parse_fasta`'s clean modes."""
import argparse
import sys

from ..fasta import parse_fasta


def build_parser():
    parser = argparse.ArgumentParser()
    parser.add_argument("-i", default=None)
    parser.add_argument("-o", default=None)
    parser.add_argument("--clean_strategy", type=str, default=None, choices=["delete", "unalign", "upper"], required=True,
                        help="delete: drop a2m insertions; upper: keep the length; unalign: drop every gap")
    parser.add_argument("--full_name", action="store_true", default=False,
                        help="if true then keep the whole name of the sequences, including the description")
    return parser


def cli(argv=None):
    args = build_parser().parse_args(argv)
    input_handle = open(args.i, "r") if args.i is not None else sys.stdin
    output_handle = open(args.o, "w") if args.o is not None else sys.stdout
    try:
        names, seqs = parse_fasta(input_handle, return_names=True, clean=args.clean_strategy, full_name=args.full_name)
        for name, seq in zip(names, seqs):
            print(f">{name}\n{seq}", file=output_handle)
    finally:
        if args.i is not None:
            input_handle.close()
        if args.o is not None:
            output_handle.close()


if __name__ == "__main__":
    cli()
