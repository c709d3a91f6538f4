"""Sample histogram over 50-instruction regions of one kernel (from an .ncu-rep with --import-source on)."""
import csv, io, subprocess, sys
rep, kernel = sys.argv[1], sys.argv[2]
raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--kernel-name", kernel, "--launch-count", "1"],
                     capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
h = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
hdr = rows[h]; ix = {c: i for i, c in enumerate(hdr)}
data = [r for r in rows[h + 1:] if len(r) == len(hdr) and r[0].startswith("0x")]
data = data[:len(data) // 2]
print("total samples", sum(int(r[ix["# Samples"]]) for r in data))
cur = 0; start = 0
for k, r in enumerate(data):
    cur += int(r[ix["# Samples"]])
    if (k + 1) % 50 == 0 or k == len(data) - 1:
        if cur >= int(sys.argv[3]) if len(sys.argv) > 3 else 20:
            best = max(data[start:k + 1], key=lambda r: int(r[ix["# Samples"]]))
            print(f"{start:5d}-{k:5d} samples {cur:5d}  exec~{data[start][ix['Instructions Executed']]:>8s}  top: "
                  f"{best[ix['Source']].strip()[:70]} ({best[ix['# Samples']]})")
        cur = 0; start = k + 1
