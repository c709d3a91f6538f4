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


def pgen_msa(templates_path, references_path, output_path, seqs_per_template, keep_identical, steps, passes, burn_in,
             device, model, alignment_size, ep, op, top_k, legacy=False, gap_percent_threshold=80, debug=False,
             sampler=None):
    template_seqs = list(zip(*parse_fasta(templates_path, clean="unalign", return_names=True)))
    reference_list = parse_fasta(references_path, clean="unalign")
    gibbs_sampler = sampler if sampler is not None else ESM_MSA_sampler(model_map[model](), device=device)

    tmp_file = tempfile.NamedTemporaryFile(delete=False, mode="w")
    write_sequential_fasta(tmp_file, reference_list)   # phmmer database: references renamed 0..n-1
    tmp_file.close()
    reference_db_path = tmp_file.name
    reference_seqs = {str(i): s for i, s in enumerate(reference_list)}
    try:
        with tqdm(total=len(template_seqs) * seqs_per_template) as pbar, open(output_path, "w") as outfile:
            for template_name, template_seq in template_seqs:
                unaligned = [template_seq]
                for hit in run_phmmer(template_seq, reference_db_path, max_mode=debug):
                    if len(unaligned) == alignment_size:
                        break
                    if reference_seqs[hit] != template_seq or keep_identical:
                        unaligned.append(reference_seqs[hit])
                if len(unaligned) < alignment_size:
                    warnings.warn(f"Warning: fewer than {alignment_size - 1} hits found for template seq {template_name}")
                _, alignment = generate_alignment({"1": unaligned}, ep=ep, op=op)   # mafft keeps the input order
                exclude_positions = []
                if not legacy:   # template is row 0: drop its gap columns, skip mostly-gap columns
                    gaps = [i for i, c in enumerate(alignment[0]) if c == "-"]
                    alignment = delete_msa_cols(alignment, gaps)
                    exclude_positions = apply_gap_threshold(alignment, gap_percent_threshold)
                else:            # original behaviour: the template is the LAST row
                    alignment[0], alignment[-1] = alignment[-1], alignment[0]
                # all seqs_per_template chains of this template as one device batch (the reference runs them one by
                # one, a batch-1 forward per step: pgen_msa_revised.py:107-115); same RNG consumption order
                new_seqs = gibbs_sampler.generate_single_batch(alignment, seqs_per_template, steps=steps, passes=passes,
                                                               burn_in=burn_in, k=top_k,
                                                               target_index=-1 if legacy else 0,
                                                               exclude_positions=exclude_positions)
                for i, new_seq in enumerate(new_seqs):
                    print(f">{i}_{template_name}\n{new_seq.replace('-', '')}", file=outfile, flush=True)
                    pbar.update(1)
    finally:
        os.unlink(reference_db_path)


def build_parser():
    parser = argparse.ArgumentParser(
        description=textwrap.dedent("""Samples from the ESM-MSA model to generate new protein sequences."""),
        formatter_class=RawAndDefaultsFormatter)
    parser.add_argument("--templates", default=None, required=True,
                        help="an unaligned fasta file with sequences to mask for generating new sequences.")
    parser.add_argument("--references", default=None, required=True,
                        help="an unaligned fasta file with reference sequences to search for homologs to the templates.")
    parser.add_argument("-o", default=None, required=True, help="a fasta file to write generated sequences to")
    parser.add_argument("--seqs_per_template", type=int, default=1,
                        help="Number of new sequences to generate for each template sequence.")
    parser.add_argument("--keep_identical", action="store_true", default=False,
                        help="By default, if a template sequence is identical to the query sequence, it is thrown out. "
                             "Set this if, for some reason you want to keep those.")
    parser.add_argument("--steps", type=int, default=10, help="Randomly assign the input positions to this many mask "
                        "bins, and mask and generate over one bin at a time.")
    parser.add_argument("--passes", type=int, default=3, help="how many passes over the entire template sequence to make.")
    parser.add_argument("--burn_in", type=int, default=1, help="A number of passes equal to burn_in will sample from the "
                        "entire distribution, after which amino acids will be sampled from the top_k most likely.")
    parser.add_argument("--top_k", type=int, default=1, help="Sample from the this many of the most probable amino "
                        "acids, after burn in. If 0 then always sample from full distribution.")
    parser.add_argument("--legacy", action="store_true", default=False,
                        help="Use the original implementation's behavior of sampling from the last sequence in the MSA "
                             "rather than the first, and ignoring gap_percent_threshold.")
    parser.add_argument("--gap_percent_threshold", type=float, default=80.0,
                        help="Don't resample positions where more than this percent of sequences in the alignment "
                             "contain gaps. Ignored in legacy mode.")
    parser.add_argument("--ep", type=float, default=0.0, help="ep parameter passed to MAFFT for alignments")
    parser.add_argument("--op", type=float, default=1.53, help="op parameter passed to MAFFT for alignments")
    parser.add_argument("--device", type=str, default="gpu", help="gpu (cuda:0) or cuda:[int]; the engine has no cpu path")
    parser.add_argument("--model", type=str, default="esm_msa1", choices=sorted(model_map), help="which model to use")
    parser.add_argument("--alignment_size", type=int, default=32, help="how many sequences (template plus references) "
                        "should be in the alignments used for sequence generation.")
    parser.add_argument("--debug", action="store_true", default=False, help="run in debug mode. Runs phmmer in --max "
                        "mode, to turn off pre-filters and allow finding very short hits.")
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
