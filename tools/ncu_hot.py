"""Top stall sites of one kernel from an .ncu-rep captured with --import-source on (read here, no GPU):
    python tools/ncu_hot.py gpurun_out/x.ncu-rep kernel_name [n]"""
import csv, io, subprocess, sys


def main(rep, kernel, n=40):
    raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--kernel-name", kernel, "--launch-count", "1"],
                         capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    h = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
    hdr = rows[h]
    ix = {c: i for i, c in enumerate(hdr)}
    data = [r for r in rows[h + 1:] if len(r) == len(hdr) and r[0].startswith("0x")]
    tot = sum(int(r[ix["# Samples"]]) for r in data)
    stalls = [c for c in hdr if c.startswith("stall_") and "Not Issued" not in c]
    print("total samples", tot, "instructions", len(data))
    agg = {s: sum(int(r[ix[s]]) for r in data) for s in stalls}
    print("stall totals:", sorted(agg.items(), key=lambda kv: -kv[1])[:8])
    for r in sorted(data, key=lambda r: -int(r[ix["# Samples"]]))[:n]:
        st = sorted(((int(r[ix[s]]), s[6:]) for s in stalls), reverse=True)[:2]
        print(r[ix["# Samples"]].rjust(6), r[ix["Instructions Executed"]].rjust(9), r[ix["Source"]].strip()[:64].ljust(64), st)


if __name__ == "__main__":
    main(sys.argv[1], sys.argv[2], int(sys.argv[3]) if len(sys.argv) > 3 else 40)
