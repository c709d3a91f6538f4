"""likelihood_esm.py drop-in (`/root/reference/src/pgen/likelihood_esm.py`): FASTA -> per-sequence mean
pseudo-log-likelihood table (tab or comma separated), optionally the position-wise values.

Scoring runs on the engine: strided <mask> copies built on the device, LM head on the masked rows only, fused
log_softmax + gather (`ESM_sampler.log_likelihood_batch` -> `Engine.score`)."""
import argparse
import sys
import textwrap

from tqdm import trange

from .. import models
from ..esm_sampler import ESM_sampler
from ..fasta import RawAndDefaultsFormatter, parse_fasta
from . import add_weight_flags, build_model

POSITIONAL_SCORE_SEP = ";"

model_map = {"esm1b": models.ESM1b, "esm6": models.ESM6, "esm12": models.ESM12, "esm34": models.ESM34,
             "esm1v": models.ESM1v, "esm2_t6_8M": models.ESM2_t6_8M, "esm2_t30_150M": models.ESM2_t30_150M,
             "esm2_t33_650M": models.ESM2_t33_650M}


def write_scores(names, scores, output_h, positionwise_h, sep):
    """One `id <sep> score` line per sequence; position-wise values rounded to 3 decimals, ';'-joined (:44-46)."""
    for name, (score, positional) in zip(names, scores):
        print(f"{name}{sep}{score}", file=output_h)
        if positionwise_h is not None:
            print(f"{name}{sep}{POSITIONAL_SCORE_SEP.join(str(round(x, 3)) for x in positional)}", file=positionwise_h)
    output_h.flush()
    if positionwise_h is not None:
        positionwise_h.flush()


def main(input_h, output_h, masking_off, sampler, batch_size, mask_distance, csv, score_name, positionwise=None,
         show_progress_bar=True):
    names, seqs = parse_fasta(input_h, return_names=True, clean="unalign")
    sep = "," if csv else "\t"
    positionwise_h = open(positionwise, "w") if positionwise is not None else None
    try:
        print(f"id{sep}{score_name}", file=output_h)
        if positionwise_h is not None:
            print(f"id{sep}{score_name}", file=positionwise_h)
        # `batch_size` sequences per call, and `batch_size` masked copies per forward inside it (:37-42)
        for i in trange(0, len(seqs), batch_size, disable=not show_progress_bar):
            scores = sampler.log_likelihood_batch(seqs[i:i + batch_size], with_masking=not masking_off,
                                                  mask_distance=mask_distance, batch_size=batch_size)
            write_scores(names[i:i + batch_size], scores, output_h, positionwise_h, sep)
    finally:
        if positionwise_h is not None:
            positionwise_h.close()


def build_parser():
    parser = argparse.ArgumentParser(
        description=textwrap.dedent("""Pseudo-log-likelihood of every sequence of a fasta under an ESM masked language
            model (B200 engine): mean over positions of log p(true residue | rest, position masked).

            Output: one line per sequence, name and score, tab (or comma) separated.
            """), formatter_class=RawAndDefaultsFormatter)
    parser.add_argument("-o", type=str, default=None, help="output table (default: stdout)")
    parser.add_argument("-i", default=None, help="fasta to score (gaps and '*' are stripped first); default stdin")
    parser.add_argument("--batch_size", type=int, default=1, help="sequences per call and masked copies per forward")
    parser.add_argument("--device", type=str, default="gpu", help="gpu (cuda:0) or cuda:[int]; the engine has no cpu path")
    parser.add_argument("--masking_off", action="store_true", default=False, help="score every position from ONE unmasked forward")
    parser.add_argument("--mask_distance", type=int, default=None,
                        help="mask every mask_distance-th position of a copy at once (mask_distance copies per sequence instead "
                             "of one per residue); default: one position per copy")
    parser.add_argument("--model", type=str, default="esm1v", choices=sorted(model_map), help="model triple (architecture + alphabet)")
    parser.add_argument("--csv", action="store_true", default=False, help="comma instead of tab separated")
    parser.add_argument("--score_name", type=str, default=None, help="header of the score column (default: the model name)")
    parser.add_argument("--positionwise", type=str, default=None,
                        help="also write per-position values here: name, then the ';'-joined list rounded to 3 decimals")
    add_weight_flags(parser)
    return parser


def mask_distance_arg(args):
    """The --mask_distance / --masking_off rules shared with likelihood_esm_msa (:85-92)."""
    mask_distance = float("inf") if args.mask_distance is None else args.mask_distance
    if mask_distance < 1:
        raise ValueError("mask distance must be an integer >= 1.")
    if args.masking_off and args.mask_distance is not None:
        raise ValueError("--masking_off and --mask_distance are both set, that doesn't make sense.")
    return mask_distance


def cli(argv=None):
    args = build_parser().parse_args(argv)
    mask_distance = mask_distance_arg(args)
    input_handle = open(args.i, "r") if args.i is not None else sys.stdin
    output_handle = open(args.o, "w") if args.o is not None else sys.stdout
    try:
        sampler = ESM_sampler(build_model(model_map, args), device=args.device)
        main(input_handle, output_handle, args.masking_off, sampler, args.batch_size, mask_distance, args.csv,
             args.score_name or args.model, args.positionwise)
    finally:
        if args.i is not None:
            input_handle.close()
        if args.o is not None:
            output_handle.close()


if __name__ == "__main__":
    cli()
