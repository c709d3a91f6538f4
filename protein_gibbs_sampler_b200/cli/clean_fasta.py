"""clean_fasta.py drop-in (`/root/reference/src/pgen/clean_fasta.py`): re-write a FASTA through one of
`parse_fasta`'s clean modes."""
import argparse
import contextlib
import sys

from ..fasta import parse_fasta

MODES = {"delete": "drop a2m insertions (lower case, '.') and '*': alignment columns only",
         "upper": "keep the length: upper-case, '.' becomes '-', '*' dropped",
         "unalign": "plain sequences: upper-case, every gap character dropped"}


def build_parser():
    parser = argparse.ArgumentParser(description="Normalise the sequences of a fasta / a2m file.")
    parser.add_argument("-i", default=None, help="input fasta (default stdin)")
    parser.add_argument("-o", default=None, help="output fasta (default stdout)")
    parser.add_argument("--clean_strategy", type=str, default=None, choices=sorted(MODES), required=True,
                        help="; ".join("%s = %s" % kv for kv in sorted(MODES.items())))
    parser.add_argument("--full_name", action="store_true", default=False,
                        help="keep the whole header line as the record name, not only its first word")
    return parser


def cli(argv=None):
    args = build_parser().parse_args(argv)
    with contextlib.ExitStack() as stack:
        src = stack.enter_context(open(args.i)) if args.i is not None else sys.stdin
        dst = stack.enter_context(open(args.o, "w")) if args.o is not None else sys.stdout
        records = parse_fasta(src, return_names=True, clean=args.clean_strategy, full_name=args.full_name)
        dst.writelines(">%s\n%s\n" % record for record in zip(*records))


if __name__ == "__main__":
    cli()
