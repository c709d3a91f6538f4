"""pgen_esm.py drop-in (`/root/reference/src/pgen/pgen_esm.py`): TSV of `name <tab> dict of sampler arguments`
-> `ESM_sampler.generate` on the GPU -> one FASTA per line plus an echo of the specification."""
import argparse
import sys
import textwrap
from pathlib import Path

from .. import models
from ..esm_sampler import ESM_sampler
from ..fasta import RawAndDefaultsFormatter, write_sequential_fasta
from . import add_weight_flags, build_model, spec_args, spec_lines

# the reference's names (pgen_esm.py:10) plus the ESM-2 family (BASELINE configs 1 and 4)
model_map = {"esm1b": models.ESM1b, "esm6": models.ESM6, "esm12": models.ESM12, "esm34": models.ESM34,
             "esm1v": models.ESM1v, "esm2_t6_8M": models.ESM2_t6_8M,
             "esm2_t30_150M": models.ESM2_t30_150M, "esm2_t33_650M": models.ESM2_t33_650M}

EPILOG = """
Keys of the sampler-argument dict (all optional except seed_seq; they are the keyword arguments of
ESM_sampler.generate):

  seed_seq               starting sequence, or a list of them (one is drawn per chain)
  max_len                output length; default = length of the (longest) seed, shorter seeds are padded with <mask>
  num_iters              Gibbs iterations per batch
  num_positions          residues resampled per chain and iteration; 0 = every eligible position
  num_positions_percent  the same as a percentage of max_len (overrides num_positions)
  in_order               True: sweep the eligible positions cyclically; False: draw them at random each iteration
  indexes                1-based positions that may change; default = everything after the leader
  leader_length          number of leading residues that are never resampled
  leader_length_percent  the same as a percentage of max_len (overrides leader_length)
  top_k                  after burn-in, sample only among the k most probable residues (0 = all 20)
  burnin                 iterations that sample from the full distribution before top_k applies
                         (0 = top_k from the start, inf = never restrict)
  temperature            logits are divided by this before sampling
  mask                   False: resample without first replacing the chosen residues by <mask>
"""


def main(input_h, output_p, args):
    sampler = ESM_sampler(build_model(model_map, args), device=args.device)
    with open(output_p / "specification.tsv", "w") as echo_h:
        for name, arg_text in spec_lines(input_h, echo_h, 2, "name, line_args"):
            sequences = sampler.generate(args.num_output_sequences, batch_size=args.batch_size, **spec_args(arg_text))
            write_sequential_fasta(output_p / (name + ".fasta"), sequences)


def build_parser():
    parser = argparse.ArgumentParser(
        description=textwrap.dedent("""Gibbs-sample new protein sequences from an ESM masked language model (B200 engine).

            One run per input line:  <run name> TAB <python dict of sampler arguments>
            Writes <run name>.fasta per line and a copy of the accepted lines to specification.tsv.
            """),
        epilog=EPILOG, formatter_class=RawAndDefaultsFormatter)
    parser.add_argument("-o", default=".", help="output directory (created if missing)")
    parser.add_argument("-i", default=None, help="specification file, one run per line: name TAB dict; default stdin")
    parser.add_argument("--batch_size", type=int, default=1, help="chains resampled together on the GPU")
    parser.add_argument("--num_output_sequences", type=int, default=1, help="sequences written per run")
    parser.add_argument("--device", type=str, default="gpu", help="gpu (cuda:0) or cuda:[int]; the engine has no cpu path")
    parser.add_argument("--model", type=str, default="esm1b", choices=sorted(model_map), help="model triple (architecture + alphabet)")
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
