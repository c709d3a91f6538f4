"""pgen_msa_revised.py drop-in (`/root/reference/src/pgen/pgen_msa_revised.py`): per template sequence, phmmer hits
from a reference set -> mafft alignment -> `ESM_MSA_sampler.generate_single` on the GPU.

phmmer and mafft stay external programs run as subprocesses (out of scope, host tools);
only the sampling runs on the engine.  The alignment-column helpers keep the reference's semantics
(`pgen_msa_revised.py:16-48`)."""
import argparse
import os
import subprocess
import sys
import tempfile
import textwrap
import warnings

from tqdm import tqdm

from .. import models
from ..esm_msa_sampler import ESM_MSA_sampler
from ..fasta import (RawAndDefaultsFormatter, parse_fasta, parse_fasta_string, write_partitioned_fasta,
                     write_sequential_fasta)
from . import add_weight_flags, build_model

model_map = {"esm_msa1": models.ESM_MSA1}


def delete_msa_cols(msa, cols):
    """The alignment without the columns whose indices are in `cols`."""
    drop = set(cols)
    return ["".join(c for i, c in enumerate(seq) if i not in drop) for seq in msa]


def count_gaps_per_column(msa):
    return [sum(seq[i] == "-" for seq in msa) for i in range(len(msa[0]))]


def apply_gap_threshold(msa, gap_threshold):
    """Indices of the columns in which MORE than gap_threshold percent of the rows hold a gap."""
    bound = len(msa) * (gap_threshold / 100)
    return [i for i, n in enumerate(count_gaps_per_column(msa)) if n > bound]


def run_phmmer(query, database, evalue=10, cpu=2, max_mode=False):
    """Names of the phmmer hits of `query` in the FASTA `database`, best first (`utils.py:265-314`).  The ranked
    hit list is read from --tblout (same order as the text report the reference parses with Biopython)."""
    with tempfile.TemporaryDirectory() as tmp:
        query_path, tbl_path = os.path.join(tmp, "query.fa"), os.path.join(tmp, "hits.tbl")
        with open(query_path, "w") as fh:
            print(f">QUERY\n{query}", file=fh)
        cmd = ["phmmer", "--noali", "--notextw", "--cpu", str(cpu), "-E", str(evalue), "--tblout", tbl_path]
        if max_mode:
            cmd.append("--max")
        out = subprocess.run(cmd + [query_path, database], stdout=subprocess.PIPE, stderr=subprocess.PIPE,
                             encoding="utf-8")
        if out.returncode != 0:
            print(f"Error in hmmer execution: \n{out.stdout}\n{out.stderr}", file=sys.stderr)
            sys.exit(1)
        with open(tbl_path) as fh:
            return [ln.split()[0] for ln in fh if ln.strip() and not ln.startswith("#")]


def generate_alignment(sequences, ep=0.0, op=1.53):
    """mafft G-INS-i alignment of {category: [seq, ...]}; (names, aligned sequences) in input order."""
    with tempfile.TemporaryDirectory() as tmp:
        path = os.path.join(tmp, "tmp.fasta")
        write_partitioned_fasta(path, sequences)
        out = subprocess.run(["mafft", "--thread", "8", "--maxiterate", "1000", "--globalpair", "--ep", str(ep),
                              "--op", str(op), path], stdout=subprocess.PIPE, stderr=subprocess.PIPE)
        if out.returncode != 0:
            print(out.stderr, file=sys.stderr)
            raise Exception("mafft failed")
    return parse_fasta_string(out.stdout.decode("utf-8"), True)


class ReferenceDb:
    """The reference sequences as a phmmer database: a temporary fasta with the records renamed 0..n-1 (the names
    phmmer reports), removed again on exit."""

    def __init__(self, sequences):
        self.by_name = {str(i): s for i, s in enumerate(sequences)}
        with tempfile.NamedTemporaryFile(delete=False, mode="w") as fh:
            write_sequential_fasta(fh, sequences)
        self.path = fh.name

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        os.unlink(self.path)

    def alignment_with_top_hits(self, name, query, size, keep_identical, ep=0.0, op=1.53, max_mode=False):
        """mafft alignment of `query` (row 0: mafft keeps the input order) with its best phmmer hits, `size` rows at
        most; hits equal to the query are skipped unless keep_identical."""
        rows = [query]
        for hit in run_phmmer(query, self.path, max_mode=max_mode):
            if len(rows) == size:
                break
            if keep_identical or self.by_name[hit] != query:
                rows.append(self.by_name[hit])
        if len(rows) < size:
            warnings.warn(f"Warning: fewer than {size - 1} hits found for template seq {name}")
        return generate_alignment({"1": rows}, ep=ep, op=op)[1]


def design_frame(alignment, legacy, gap_percent_threshold):
    """(alignment to sample on, row to resample, 0-based columns to leave alone).  Default: the template stays row 0,
    its gap columns are removed and mostly-gap columns are excluded.  Legacy: the template is swapped to the LAST row
    and every column may change."""
    if legacy:
        return [alignment[-1]] + alignment[1:-1] + [alignment[0]] if len(alignment) > 1 else list(alignment), -1, []
    trimmed = delete_msa_cols(alignment, [i for i, c in enumerate(alignment[0]) if c == "-"])
    return trimmed, 0, apply_gap_threshold(trimmed, gap_percent_threshold)


def pgen_msa(templates_path, references_path, output_path, seqs_per_template, keep_identical, steps, passes, burn_in,
             device, model, alignment_size, ep, op, top_k, legacy=False, gap_percent_threshold=80, debug=False,
             sampler=None):
    names, templates = parse_fasta(templates_path, clean="unalign", return_names=True)
    gibbs_sampler = sampler if sampler is not None else ESM_MSA_sampler(model_map[model](), device=device)
    with ReferenceDb(parse_fasta(references_path, clean="unalign")) as db, open(output_path, "w") as outfile, \
            tqdm(total=len(templates) * seqs_per_template) as pbar:
        for name, template in zip(names, templates):
            alignment = db.alignment_with_top_hits(name, template, alignment_size, keep_identical, ep, op, debug)
            alignment, target, frozen = design_frame(alignment, legacy, gap_percent_threshold)
            # all seqs_per_template chains of this template as ONE device batch (the reference runs them one by one,
            # a batch-1 forward per step: pgen_msa_revised.py:107-115); the RNG is consumed in the same order
            designs = gibbs_sampler.generate_single_batch(alignment, seqs_per_template, steps=steps, passes=passes,
                                                          burn_in=burn_in, k=top_k, target_index=target,
                                                          exclude_positions=frozen)
            for k, design in enumerate(designs):
                outfile.write(">%d_%s\n%s\n" % (k, name, design.replace("-", "")))
                outfile.flush()
                pbar.update(1)


def build_parser():
    parser = argparse.ArgumentParser(
        description=textwrap.dedent("""Samples from the ESM-MSA model to generate new protein sequences."""),
        formatter_class=RawAndDefaultsFormatter)
    parser.add_argument("--templates", default=None, required=True,
                        help="fasta of the (unaligned) sequences to re-design, one after the other")
    parser.add_argument("--references", default=None, required=True,
                        help="fasta of (unaligned) sequences searched with phmmer for homologs of each template")
    parser.add_argument("-o", default=None, required=True, help="output fasta; records are named <k>_<template name>")
    parser.add_argument("--seqs_per_template", type=int, default=1,
                        help="designs produced per template (run as one device batch)")
    parser.add_argument("--keep_identical", action="store_true", default=False,
                        help="keep phmmer hits that are identical to the template (dropped otherwise)")
    parser.add_argument("--steps", type=int, default=10, help="each pass shuffles the positions into this many "
                        "bins; one bin is masked and resampled per forward")
    parser.add_argument("--passes", type=int, default=3, help="number of sweeps over all positions")
    parser.add_argument("--burn_in", type=int, default=1, help="sweeps that sample from the full "
                        "distribution before --top_k takes effect")
    parser.add_argument("--top_k", type=int, default=1, help="after burn-in draw only among the k likeliest "
                        "residues (0: never restrict)")
    parser.add_argument("--legacy", action="store_true", default=False,
                        help="first version's behaviour: the template goes LAST in the alignment and is resampled there; "
                             "--gap_percent_threshold is not applied")
    parser.add_argument("--gap_percent_threshold", type=float, default=80.0,
                        help="columns with more than this percentage of gaps are left as they are (not in --legacy mode)")
    parser.add_argument("--ep", type=float, default=0.0, help="mafft --ep (offset / gap extension)")
    parser.add_argument("--op", type=float, default=1.53, help="mafft --op (gap opening)")
    parser.add_argument("--device", type=str, default="gpu", help="gpu (cuda:0) or cuda:[int]; the engine has no cpu path")
    parser.add_argument("--model", type=str, default="esm_msa1", choices=sorted(model_map), help="model triple (architecture + alphabet)")
    parser.add_argument("--alignment_size", type=int, default=32, help="rows of each working alignment: the template "
                        "plus its best hits")
    parser.add_argument("--debug", action="store_true", default=False, help="phmmer --max (no pre-filters: finds "
                        "hits of very short sequences too)")
    add_weight_flags(parser)
    return parser


def main(argv):
    args = build_parser().parse_args(argv)
    sampler = ESM_MSA_sampler(build_model(model_map, args), device=args.device)
    pgen_msa(args.templates, args.references, args.o, args.seqs_per_template, args.keep_identical, args.steps,
             args.passes, args.burn_in, args.device, args.model, args.alignment_size, args.ep, args.op, args.top_k,
             legacy=args.legacy, gap_percent_threshold=args.gap_percent_threshold, debug=args.debug, sampler=sampler)


if __name__ == "__main__":
    main(sys.argv[1:])
