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
                                             ("pgen_msa_revised", "pgen_msa_revised.py"),
                                             ("pgen_esm_from_fasta", "pgen_esm_from_fasta.py"),
                                             ("likelihood_esm", "likelihood_esm.py"),
                                             ("likelihood_esm_msa", "likelihood_esm_msa.py"),
                                             ("clean_fasta", "clean_fasta.py")])
def test_cli_flags_match_reference(module, ref_file):
    """Every flag of the reference script exists here with the same default and action; only --device defaults to
    the GPU (the engine has no CPU path) and --model lists the models this engine implements."""
    import importlib
    mod = importlib.import_module("protein_gibbs_sampler_b200.cli." + module)
    ours = _flags(mod.build_parser())
    path = os.path.join(REF, ref_file)
    if not os.path.exists(path):
        pytest.skip("reference source not present on this machine")
    src = "".join(ln for ln in open(path) if not ln.lstrip().startswith("#"))   # commented-out flags do not exist
    src = re.sub(r"add_argument\(\s*'(-{1,2}[A-Za-z_]+)'", r'add_argument("\1"', src)   # either quote style
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
    assert "--device" not in ours or ours["--device"][0] == "gpu"


def test_clean_fasta_cli(tmp_path):
    from protein_gibbs_sampler_b200.cli import clean_fasta
    (tmp_path / "in.a2m").write_text(A2M)
    for mode, want in (("delete", CORE), ("unalign", FULL)):
        clean_fasta.cli(["-i", str(tmp_path / "in.a2m"), "-o", str(tmp_path / "out.fa"), "--clean_strategy", mode])
        names, seqs = fasta.parse_fasta(tmp_path / "out.fa", return_names=True)
        assert names == ["seq_1", "seq_2", "seq_3"] and seqs[0] == want and seqs[2] == want
    clean_fasta.cli(["-i", str(tmp_path / "in.a2m"), "-o", str(tmp_path / "out.fa"), "--clean_strategy", "upper",
                     "--full_name"])
    assert fasta.parse_fasta(tmp_path / "out.fa", return_names=True, full_name=True)[0][1] == \
        "seq_2 second record has a description"


def _loglik_golden():
    import json
    with open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "reference_loglik.json")) as f:
        return json.load(f)


def test_likelihood_esm_cli_table_vs_reference_values(tmp_path):
    """Host logic of the likelihood_esm drop-in: FASTA in (gaps / stop codons stripped), `id <sep> score` table and the
    ';'-joined position-wise file out; the numbers are the reference sampler's own (golden, fp32 oracle model)."""
    from oracle.fair_esm import OracleModel
    from protein_gibbs_sampler_b200.cli import likelihood_esm
    from protein_gibbs_sampler_b200.esm_sampler import ESM_sampler
    from protein_gibbs_sampler_b200.weights import synthetic_state_dict
    c = [c for c in _loglik_golden()["esm"] if len(c["seqs"]) == 2 and c["kwargs"].get("mask_distance") == 4][0]
    s = ESM_sampler(OracleModel(c["cfg"], synthetic_state_dict(c["cfg"], c["weights_seed"])), device="cpu")
    text = ">first some description\n%s-*\n>second\n%s\n" % (c["seqs"][0][:4] + "-" + c["seqs"][0][4:].lower(), c["seqs"][1])
    for csv, sep in ((False, "\t"), (True, ",")):
        out, pos = io.StringIO(), tmp_path / "pos.txt"
        likelihood_esm.main(io.StringIO(text), out, False, s, 2, 4, csv, "myscore", str(pos), show_progress_bar=False)
        lines = out.getvalue().strip().split("\n")
        assert lines[0] == "id%smyscore" % sep and [ln.split(sep)[0] for ln in lines[1:]] == ["first", "second"]
        for ln, (mean, each) in zip(lines[1:], c["result"]):
            assert float(ln.split(sep)[1]) == pytest.approx(mean, abs=2e-6)
        plines = pos.read_text().strip().split("\n")
        assert plines[0] == "id%smyscore" % sep
        for ln, (mean, each) in zip(plines[1:], c["result"]):
            assert [float(v) for v in ln.split(sep)[1].split(";")] == pytest.approx([round(v, 3) for v in each], abs=1.1e-3)
    args = likelihood_esm.build_parser().parse_args(["--masking_off", "--mask_distance", "3"])
    with pytest.raises(ValueError, match="both set"):
        likelihood_esm.mask_distance_arg(args)
    with pytest.raises(ValueError, match=">= 1"):
        likelihood_esm.mask_distance_arg(likelihood_esm.build_parser().parse_args(["--mask_distance", "0"]))


def test_likelihood_esm_msa_cli_contexts(tmp_path):
    """likelihood_esm_msa drop-in, host logic: the query goes on top of the (subset of the) reference alignment,
    columns where the query has a gap are deleted, the top row is scored; `in_msas` bypasses the reference MSA."""
    from oracle.fair_esm import OracleModel
    from protein_gibbs_sampler_b200.cli import likelihood_esm_msa as cli
    from protein_gibbs_sampler_b200.esm_msa_sampler import ESM_MSA_sampler
    from protein_gibbs_sampler_b200.weights import synthetic_state_dict
    c = _loglik_golden()["msa"][0]
    s = ESM_MSA_sampler(OracleModel(c["cfg"], synthetic_state_dict(c["cfg"], c["weights_seed"])), device="cpu")
    ref = ["MKTAYIAK-RQ", "MKTAYLAKQRQ", "MRTAY-AKQRQ"]
    queries = {"q1": "MK-AYIAKQRQ", "q2": "MKTAWIAKQR-"}
    qtext = "".join(">%s\n%s\n" % kv for kv in queries.items())
    rtext = "".join(">r%d\n%s\n" % (i, r) for i, r in enumerate(ref))
    out = io.StringIO()
    cli.main(io.StringIO(qtext), out, False, s, io.StringIO(rtext), subset_strategy="in_order", alignment_size=2,
             mask_distance=3, positionwise=str(tmp_path / "p.tsv"), show_progress_bar=False)
    lines = out.getvalue().strip().split("\n")
    assert lines[0] == "id\tesm-msa" and [ln.split("\t")[0] for ln in lines[1:]] == ["q1", "q2"]
    for ln, (name, q) in zip(lines[1:], queries.items()):
        gaps = [i for i, ch in enumerate(q) if ch == "-"]
        msa = cli.delete_msa_cols([q] + ref[:2], gaps)
        assert all(len(r) == 10 for r in msa) and "-" not in msa[0]
        want, each = s.log_likelihood(msa, mask_distance=3)
        assert float(ln.split("\t")[1]) == pytest.approx(want, abs=1e-6) and len(each) == 10
    assert len((tmp_path / "p.tsv").read_text().strip().split("\n")[1].split("\t")[1].split(";")) == 10
    # user-supplied alignments: the golden MSA itself, query row first
    out = io.StringIO()
    cli.main(io.StringIO(">x\n%s\n" % c["msas"][0][0]), out, False, s, in_msas={"x": c["msas"][0]}, csv=True,
             show_progress_bar=False)
    gaps = [i for i, ch in enumerate(c["msas"][0][0]) if ch == "-"]
    want = s.log_likelihood(cli.delete_msa_cols(c["msas"][0], gaps))[0]
    assert float(out.getvalue().strip().split("\n")[1].split(",")[1]) == pytest.approx(want, abs=1e-6)
    # redraw advances the subset seed by 1e6 per query (reference :98-101)
    b = cli.ContextBuilder(queries, None, ref, "random", 2, subset_random_seed=5, redraw=True)
    b("q1"); b("q2")
    assert b.seed == 2000005


def test_reference_db_top_hits_and_design_frame(monkeypatch):
    """Host logic shared by pgen_msa_revised and likelihood_esm_msa --subset_strategy top_hits (phmmer / mafft stubbed:
    hits in database order, sequences right-padded with gaps): hit selection, the size cap, identical-hit filtering,
    the too-few-hits warning, clean-up of the temporary database; and the two template frames of pgen_msa_revised."""
    from protein_gibbs_sampler_b200.cli import pgen_msa_revised as cli
    seen = {}

    def fake_phmmer(query, db, max_mode=False):
        seen["db"] = db
        assert fasta.parse_fasta(db) == refs           # the database holds the references, renamed 0..n-1
        return fasta.parse_fasta(db, return_names=True)[0]

    def fake_mafft(groups, ep=0.0, op=1.53):
        rows = groups["1"]
        width = max(len(r) for r in rows)
        return ["1_%d" % i for i in range(len(rows))], [r + "-" * (width - len(r)) for r in rows]
    monkeypatch.setattr(cli, "run_phmmer", fake_phmmer)
    monkeypatch.setattr(cli, "generate_alignment", fake_mafft)
    refs = ["MAGIK", "MEADAL", "MAGIC", "MEADQLK"]
    with cli.ReferenceDb(refs) as db:
        assert db.alignment_with_top_hits("t", "MAGIC", 3, keep_identical=False) == ["MAGIC-", "MAGIK-", "MEADAL"]
        assert db.alignment_with_top_hits("t", "MAGIC", 4, keep_identical=True) == \
            ["MAGIC-", "MAGIK-", "MEADAL", "MAGIC-"]
        with pytest.warns(UserWarning, match="fewer than 5 hits"):
            assert len(db.alignment_with_top_hits("t", "MAGIC", 6, keep_identical=False)) == 4
        assert os.path.exists(seen["db"])
    assert not os.path.exists(seen["db"])
    msa = ["MA-GIC", "MAAG-C", "-AAGIC"]
    assert cli.design_frame(msa, legacy=False, gap_percent_threshold=30) == (["MAGIC", "MAG-C", "-AGIC"], 0, [0, 3])
    assert cli.design_frame(msa, legacy=True, gap_percent_threshold=30) == (["-AAGIC", "MAAG-C", "MA-GIC"], -1, [])
    assert cli.design_frame(msa[:1], legacy=True, gap_percent_threshold=30) == (["MA-GIC"], -1, [])


# ------------------------------------------------------------------------------------------------ GPU: end to end
AA = set("ACDEFGHIKLMNPQRSTVWY")


@pytest.mark.gpu
def test_likelihood_cli_on_engine(tmp_path, gpu_lib):
    """Both likelihood drop-ins end to end on the engine (synthetic weights): tables well-formed, values equal to
    the sampler API's."""
    from protein_gibbs_sampler_b200.cli import likelihood_esm, likelihood_esm_msa
    seqs = {"a": "MKTAYIAKQRQISFVKSHFSRQLEE", "b": "MKTAYIAKQR"}
    (tmp_path / "in.fa").write_text("".join(">%s\n%s\n" % kv for kv in seqs.items()))
    likelihood_esm.cli(["-i", str(tmp_path / "in.fa"), "-o", str(tmp_path / "out.tsv"), "--model", "esm2_t6_8M",
                        "--device", "cuda:0", "--mask_distance", "5", "--batch_size", "2",
                        "--positionwise", str(tmp_path / "pos.tsv")])
    rows = [ln.split("\t") for ln in (tmp_path / "out.tsv").read_text().strip().split("\n")]
    assert rows[0] == ["id", "esm2_t6_8M"] and [r[0] for r in rows[1:]] == ["a", "b"]
    assert all(-8.0 < float(r[1]) < 0.0 for r in rows[1:])
    pos = [ln.split("\t") for ln in (tmp_path / "pos.tsv").read_text().strip().split("\n")][1:]
    assert [len(p[1].split(";")) for p in pos] == [25, 10]
    (tmp_path / "ref.fa").write_text(">r0\nMKTAYLAKQR\n>r1\nMRTAY-AKQR\n>r2\nMKTAWIAKQR\n")
    (tmp_path / "q.fa").write_text(">q\nMKTAYIAK-R\n")
    likelihood_esm_msa.cli(["-i", str(tmp_path / "q.fa"), "-o", str(tmp_path / "m.csv"), "--csv", "--reference_msa",
                            str(tmp_path / "ref.fa"), "--subset_strategy", "in_order", "--device", "gpu"])
    rows = [ln.split(",") for ln in (tmp_path / "m.csv").read_text().strip().split("\n")]
    assert rows[0] == ["id", "esm-msa"] and rows[1][0] == "q" and -8.0 < float(rows[1][1]) < 0.0


@pytest.mark.gpu
def test_pgen_esm_from_fasta_cli_and_generate_many(tmp_path, gpu_lib):
    """pgen_esm_from_fasta drop-in.  generate_many folds the reference's one-chain-per-output loop
    (pgen_esm_from_fasta.py:27-33) into one device batch per seed length; in replay mode the output is identical to
    that loop run call by call with the same seeds of `random` / torch."""
    import random
    import torch
    from protein_gibbs_sampler_b200 import models
    from protein_gibbs_sampler_b200.cli import pgen_esm_from_fasta as cli
    from protein_gibbs_sampler_b200.esm_sampler import ESM_sampler
    seeds = ["MKTAYIAKQR-ISFVK", "MKT.AYLAKQRQISFVKSH", "mrtayiakqrq-sfvk"]
    (tmp_path / "seeds.fa").write_text("".join(">s%d\n%s\n" % (i, q) for i, q in enumerate(seeds)))
    kw = dict(num_iters=3, burnin=1, top_k=2, num_positions_percent=30, leader_length=2)
    s = ESM_sampler(models.ESM2_t6_8M(seed=1), device="cuda:0", rng="replay")
    random.seed(3); torch.manual_seed(3)
    folded = cli.sample_from_seeds(s, seeds, 7, kw, keep_gap_positions=True)
    random.seed(3); torch.manual_seed(3)
    looped = []
    for _ in range(7):
        seed, gap_mask = fasta.unalign(random.choice(seeds))
        looped.append(fasta.add_gaps_back(s.generate(1, seed, batch_size=1, show_progress_bar=False, **kw)[0], gap_mask))
    assert folded == looped
    assert {len(q) for q in folded} <= {len(q) for q in seeds}
    spec = tmp_path / "spec.tsv"
    spec.write_text("fam\t%r\t%s\nignored line\n" % (kw, tmp_path / "seeds.fa"))
    cli.cli(["-i", str(spec), "-o", str(tmp_path / "out"), "--num_output_sequences", "5", "--device", "cuda:0",
             "--model", "esm2_t6_8M"])
    names, seqs = fasta.parse_fasta(tmp_path / "out" / "fam.fasta", return_names=True)
    assert names == [str(i) for i in range(5)] and all(set(q) <= AA and len(q) in (15, 18) for q in seqs)


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
