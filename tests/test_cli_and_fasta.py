"""Command-line drop-ins and their FASTA helpers.

CPU part: the text helpers against the values the reference's own tests pin (`/root/reference/test/test_utils.py:26-110`,
restated here, not imported: `pgen.utils` needs Biopython), the alignment-column helpers, and flag-for-flag parity of
the three argument parsers with the reference scripts (checked against the reference SOURCE when it is present).
GPU part: the CLIs end to end (BASELINE config 1 plumbing on the engine)."""
import io
import os
import re
import sys

import pytest

from protein_gibbs_sampler_b200 import fasta
from protein_gibbs_sampler_b200.cli import spec_args

REF = "/root/reference/src/pgen"

A2M = """
>seq_1
mdgtrtsldieeysdtevqknqvlTLEEWQDKWVNGKTAFHQEQGHQLLKKHLDTflKGKSGLRVFFPLCGKAVEMKWFADRGHSVVGVEISELGIQEFFTEQNLSYSeep*
>seq_2 second record has a description
........................TLEEWQDKWVNGKTAFHQEQGHQLLKKHLDT..KGKSGLRVFFPLCGKAVEMKWFADRGHSVVGVEISELGIQEFFTEQNLSYS...*
>seq_3
mdgtrtsldieeysdtevqknqvlTLEEWQDKWVNGK
TAFHQEQGHQLLKKHLDTflKGKSGLRVFFPLCGKAV
EMKWFADRGHSVVGVEISELGIQEFFTEQNLSYSeep*

"""
CORE = "TLEEWQDKWVNGKTAFHQEQGHQLLKKHLDTKGKSGLRVFFPLCGKAVEMKWFADRGHSVVGVEISELGIQEFFTEQNLSYS"
FULL = "MDGTRTSLDIEEYSDTEVQKNQVLTLEEWQDKWVNGKTAFHQEQGHQLLKKHLDTFLKGKSGLRVFFPLCGKAVEMKWFADRGHSVVGVEISELGIQEFFTEQNLSYSEEP"


@pytest.mark.parametrize("kw,expected", [
    (dict(n=1, keep_first=True, strategy="in_order"), [0]), (dict(n=1, keep_first=False, strategy="in_order"), [0]),
    (dict(n=1, keep_first=True, strategy="random"), [0]), (dict(n=3, keep_first=True, strategy="in_order"), [0, 1, 2]),
    (dict(n=3, keep_first=False, strategy="in_order"), [0, 1, 2]), (dict(n=0, keep_first=False, strategy="in_order"), []),
    (dict(n=0, keep_first=True, strategy="in_order"), []),
    (dict(n=5000, keep_first=True, strategy="in_order"), [0, 1, 2, 3, 4, 5]),
    (dict(n=5000, keep_first=False, strategy="in_order"), [0, 1, 2, 3, 4, 5])])
def test_subsetter(kw, expected):
    assert fasta.SequenceSubsetter.subset(seq_list=[0, 1, 2, 3, 4, 5], **kw) == expected


def test_subsetter_random_and_errors():
    import random
    src = [0, 1, 2, 3, 4, 5]
    out = fasta.SequenceSubsetter.subset(src, 5000, keep_first=True, strategy="random", random_seed=1)
    assert out[0] == 0 and set(out) == set(src) and out != src
    # the draw is `random.Random(seed).shuffle` of the candidates (reference utils.py:350), independent of global state
    want = src[1:]
    random.Random(1).shuffle(want)
    assert out == [0] + want
    assert src == [0, 1, 2, 3, 4, 5]
    with pytest.raises(ValueError):
        fasta.SequenceSubsetter.subset(src, 2, strategy="best")


def test_parse_fasta_clean_modes():
    names, seqs = fasta.parse_fasta(io.StringIO(A2M), return_names=True)
    assert names == ["seq_1", "seq_2", "seq_3"]
    assert seqs[0] == seqs[2] and seqs[0].endswith("eep*") and seqs[1].startswith("." * 24)
    assert fasta.parse_fasta(io.StringIO(A2M), return_names=True, full_name=True)[0][1] == \
        "seq_2 second record has a description"
    assert fasta.parse_fasta(io.StringIO(A2M), clean="delete") == [CORE] * 3
    up = fasta.parse_fasta(io.StringIO(A2M), clean="upper")
    assert up[0] == FULL and up[2] == FULL
    assert up[1] == "-" * 24 + CORE[:31] + "--" + CORE[31:] + "---" and len(up[1]) == len(FULL)
    assert fasta.parse_fasta(io.StringIO(A2M), clean="unalign") == [FULL, CORE, FULL]
    with pytest.raises(ValueError):
        fasta.parse_fasta(io.StringIO(A2M), clean="lower")


def test_fasta_roundtrip_through_files(tmp_path):
    p = tmp_path / "x.fasta"
    fasta.write_sequential_fasta(p, ["MKV", "AC-D"])
    assert p.read_text() == ">0\nMKV\n>1\nAC-D\n"
    assert fasta.parse_fasta(str(p), return_names=True) == (["0", "1"], ["MKV", "AC-D"])
    with open(p) as fh:                       # open handles are accepted and left open
        assert fasta.parse_fasta(fh) == ["MKV", "AC-D"] and not fh.closed
    fasta.write_partitioned_fasta(p, {"a": ["MK"], "b": ["AC", "DE"]})
    assert fasta.parse_fasta(p, return_names=True) == (["a_0", "b_0", "b_1"], ["MK", "AC", "DE"])


@pytest.mark.parametrize("text,letters,mask", [
    (".*-ABCDE.*-", "ABCDE", [".", "*", "-", None, None, None, None, None, ".", "*", "-"]),
    ("AB.*-ab", "ABAB", [None, None, ".", "*", "-", None, None])])
def test_unalign_and_back(text, letters, mask):
    assert fasta.unalign(text) == (letters, mask)
    assert fasta.add_gaps_back(letters, mask) == text.upper()
    assert fasta.add_gaps_back("MTGQ", [None, "-", "-", None, None, ".", "-", None, "*"]) == "M--TG.-Q*"


def test_alignment_column_helpers():
    from protein_gibbs_sampler_b200.cli.pgen_msa_revised import apply_gap_threshold, count_gaps_per_column, delete_msa_cols
    msa = ["A-C-E", "AB--E", "-BC-E", "ABC-E"]
    assert count_gaps_per_column(msa) == [1, 1, 1, 4, 0]
    assert apply_gap_threshold(msa, 80) == [3]          # strictly more than 80 % of 4 rows
    assert apply_gap_threshold(msa, 25) == [3]          # 1 gap of 4 rows is not MORE than 25 %
    assert apply_gap_threshold(msa, 24) == [0, 1, 2, 3]
    assert delete_msa_cols(msa, [3, 0]) == ["-CE", "B-E", "BCE", "BCE"]
    assert delete_msa_cols(msa, iter(())) == msa


def test_spec_args_grammar():
    """The dict literals the reference `eval`s (pgen_esm.py:25), including its example lines and inf burn-in."""
    a = spec_args("{'num_iters': 20, 'burnin': 10, 'mask': True, 'in_order':False, 'num_positions_percent': 10, "
                  "'seed_seq': 'mrhgdissSND'}")
    assert a == {"num_iters": 20, "burnin": 10, "mask": True, "in_order": False, "num_positions_percent": 10,
                 "seed_seq": "mrhgdissSND"}
    assert spec_args("{'burnin': float('inf'), 'indexes': list(range(1, 4)), 'temperature': None}") == \
        {"burnin": float("inf"), "indexes": [1, 2, 3], "temperature": None}
    with pytest.raises(Exception):
        spec_args("__import__('os').system('true')")
    with pytest.raises(ValueError):
        spec_args("[1, 2]")
    for ex in ("esm_specification.tsv", "msa_specification.tsv"):
        path = os.path.join("/root/reference/examples", ex)
        if os.path.exists(path):
            for line in open(path):
                if line.strip():
                    assert isinstance(spec_args(line.rstrip("\n").split("\t")[1]), dict)


def _flags(parser):
    return {a.option_strings[0]: (a.default, a.type, a.nargs, type(a).__name__) for a in parser._actions
            if a.option_strings and a.option_strings[0] not in ("-h", "--checkpoint", "--weights_seed")}


@pytest.mark.parametrize("module,ref_file", [("pgen_esm", "pgen_esm.py"), ("pgen_msa", "pgen_msa.py"),
                                             ("pgen_msa_revised", "pgen_msa_revised.py")])
def test_cli_flags_match_reference(module, ref_file):
    """Every flag of the reference script exists here with the same default and action; only --device defaults to
    the GPU (the engine has no CPU path) and --model lists the models this engine implements."""
    import importlib
    mod = importlib.import_module("protein_gibbs_sampler_b200.cli." + module)
    ours = _flags(mod.build_parser())
    path = os.path.join(REF, ref_file)
    if not os.path.exists(path):
        pytest.skip("reference source not present on this machine")
    src = open(path).read()
    ref_flags = re.findall(r'add_argument\(\s*"(-{1,2}[A-Za-z_]+)"', src)
    assert ref_flags and set(ref_flags) == set(ours), (sorted(ref_flags), sorted(ours))
    for flag in ref_flags:
        call = src[src.index('add_argument("%s"' % flag):]
        call = call[:call.index("\n    parser.add_argument") if "\n    parser.add_argument" in call else call.index("args = parser")]
        m = re.search(r"default=([^,\)]+)", call)
        default, typ, nargs, action = ours[flag]
        if flag in ("--device", "--model") or m is None:
            continue
        want = eval(m.group(1), {"sys": sys})
        assert default == want, (flag, default, want)
        assert ("store_true" in call) == (action == "_StoreTrueAction"), flag
    assert ours["--device"][0] == "gpu"


# ------------------------------------------------------------------------------------------------ GPU: end to end
AA = set("ACDEFGHIKLMNPQRSTVWY")


@pytest.mark.gpu
def test_pgen_esm_cli_config1(tmp_path, gpu_lib):
    """BASELINE configs[0] plumbing: esm2_t6_8M, one seed of L=25, batch 2, 5 Gibbs iterations, through the CLI."""
    from protein_gibbs_sampler_b200.cli import pgen_esm
    seed = "MKTAYIAKQRQISFVKSHFSRQLEE"
    spec = tmp_path / "spec.tsv"
    spec.write_text("run1\t{'num_iters': 5, 'burnin': 2, 'top_k': 3, 'num_positions_percent': 10, 'seed_seq': '%s'}\n"
                    "\n"
                    "bad line without a dict\n"
                    "run2\t{'num_iters': 2, 'in_order': True, 'num_positions': 4, 'leader_length': 3, 'mask': False, "
                    "'seed_seq': '%s'}\n" % (seed, seed.lower()))
    out = tmp_path / "out"
    pgen_esm.cli(["-i", str(spec), "-o", str(out), "--batch_size", "2", "--num_output_sequences", "3",
                  "--device", "cuda:0", "--model", "esm2_t6_8M"])
    assert sorted(os.listdir(out)) == ["run1.fasta", "run2.fasta", "specification.tsv"]
    assert len((out / "specification.tsv").read_text().strip().split("\n")) == 2
    for name in ("run1", "run2"):
        names, seqs = fasta.parse_fasta(out / (name + ".fasta"), return_names=True)
        assert names == ["0", "1", "2"] and all(len(s) == 25 and set(s) <= AA for s in seqs)
    # leader_length=3 with in-order sampling of 4 positions x 2 iterations from an unmasked seed: the leader and
    # every position the schedule never reached are the seed's
    seqs = fasta.parse_fasta(out / "run2.fasta")
    assert all(s[:3] == seed[:3] and s[14:] == seed[14:] for s in seqs)


@pytest.mark.gpu
def test_pgen_msa_cli(tmp_path, gpu_lib):
    from protein_gibbs_sampler_b200.cli import pgen_msa
    msa = ["MKTAYIAKQR-ISFVK", "MKTAYLAKQRQISFVK", "MRTAYIAKQRQ-SFVK", "mKTAWIAKQRQISFvK", "MKTAYIAKQRQISF.K"]
    fa = tmp_path / "seed.a2m"
    fa.write_text("".join(">s%d\n%s\n" % (i, s) for i, s in enumerate(msa)))
    spec = tmp_path / "spec.tsv"
    spec.write_text("fam\t{'num_iters': 3, 'burnin': 1, 'top_k': 2, 'num_positions_percent': 20}\t%s\n" % fa)
    out = tmp_path / "out"
    pgen_msa.cli(["-i", str(spec), "-o", str(out), "--num_output_sequences", "7", "--device", "gpu",
                  "--alignment_size", "4", "--keep_first_sequence", "--subset_strategy", "in_order"])
    names, seqs = fasta.parse_fasta(out / "fam.fasta", return_names=True)
    assert names == [str(i) for i in range(7)]
    assert all(len(s) == 16 and set(s) <= AA | {"-"} for s in seqs)


@pytest.mark.gpu
def test_pgen_msa_revised_cli_with_stubbed_tools(tmp_path, gpu_lib, monkeypatch):
    """phmmer / mafft are absent here: stub them (hits in database order, sequences right-padded with gaps) and check
    the reference test's expectations (`/root/reference/test/test_pgen_msa_revised.py:10-34`): record names
    `<i>_<template>` and output lengths bounded by the template lengths, in both legacy and default mode."""
    from protein_gibbs_sampler_b200.cli import pgen_msa_revised as cli
    (tmp_path / "q.fasta").write_text(">query_seq1 first\nMAGIC\n>query_seq2\nMEADAL\n")
    (tmp_path / "r.fasta").write_text(">r0\nMAGIK\n>r1\nMEADAL\n>r2\nMEADQL\n>r3\nMAGI\n>r4\nMEADALQ\n")
    monkeypatch.setattr(cli, "run_phmmer", lambda query, db, max_mode=False: fasta.parse_fasta(db, return_names=True)[0])

    def fake_mafft(groups, ep=0.0, op=1.53):
        seqs = groups["1"]
        width = max(len(s) for s in seqs)
        return ["1_%d" % i for i in range(len(seqs))], [s + "-" * (width - len(s)) for s in seqs]
    monkeypatch.setattr(cli, "generate_alignment", fake_mafft)
    for extra in (["--legacy", "--alignment_size", "1"], ["--alignment_size", "5", "--gap_percent_threshold", "49", "--debug"]):
        out = tmp_path / "gen.fasta"
        cli.main(["--templates", str(tmp_path / "q.fasta"), "--references", str(tmp_path / "r.fasta"), "-o", str(out),
                  "--seqs_per_template", "2", "--steps", "2", "--passes", "2", "--device", "cuda:0"] + extra)
        names, seqs = fasta.parse_fasta(out, return_names=True)
        assert names == ["0_query_seq1", "1_query_seq1", "0_query_seq2", "1_query_seq2"]
        # The reference test expects lengths 5,5,6,6 with pretrained weights; the MSA sampler may draw '-' (it is in
        # its candidate set, and the CLI strips it), which random weights do now and then, so: at most that long.
        assert all(1 <= len(s) <= n and set(s) <= AA for s, n in zip(seqs, [5, 5, 6, 6])), seqs
