"""SASS evidence table per kernel of the built library (no GPU needed):
    python tools/sass_evidence.py protein_gibbs_sampler_b200/libpgibbs.so profiles/x_sass_evidence.txt"""
import collections, re, subprocess, sys

COLS = ["UTCHMMA", "UTCBAR", "UTMALDG", "UTMASTG", "UTMAREDG", "LDTM", "STTM", "HMMA", "MUFU.EX2", "FFMA2", "FMUL2", "LDGSTS", "SYNCS",
        "LDL", "STL"]


def main(lib, out):
    sass = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True).stdout
    name, acc = None, collections.OrderedDict()
    for line in sass.splitlines():
        m = re.search(r"Function : (\S+)", line)
        if m:
            name = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip()
            name = re.sub(r"\(.*", "", name).replace("void ", "").replace("pg::", "")
            acc[name] = collections.Counter()
            continue
        m = re.match(r"\s+/\*[0-9a-f]+\*/\s+(?:@!?U?P\d\s+)?([A-Z0-9_.]+)", line)
        if m and name:
            op = m.group(1)
            acc[name]["instrs"] += 1
            for c in COLS:
                if op == c or op.startswith(c + "."):
                    acc[name][c] += 1
    with open(out, "w") as f:
        f.write("# SASS evidence per kernel of %s (cuobjdump -sass; sm_100a).\n" % lib)
        f.write("# UTCHMMA = tcgen05.mma, UTCBAR = tcgen05.commit, UTMALDG / UTMASTG / UTMAREDG = TMA load / store / reduce-add,\n"
                "# LDTM / STTM = tcgen05.ld / st, SYNCS = mbarrier ops, HMMA = legacy mma.sync, LDGSTS = cp.async, FFMA2 / FMUL2 = packed\n"
                "# fp32 pairs, LDL / STL = local-memory (spill) accesses.\n")
        f.write("%-66s %7s " % ("kernel", "instrs") + " ".join("%8s" % c for c in COLS) + "\n")
        for k, v in acc.items():
            f.write("%-66s %7d " % (k[:66], v["instrs"]) + " ".join("%8d" % v[c] for c in COLS) + "\n")
    print(open(out).read()[:1500])


if __name__ == "__main__":
    main(sys.argv[1], sys.argv[2])
