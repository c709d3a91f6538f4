"""Summarise an ncu launch list (--metrics gpu__time_duration.sum --csv) per kernel: launches, total and mean time,
share of the listed time.    python tools/launch_list_summary.py gpurun_out/x_launches.csv profiles/x_launches_summary.txt"""
import collections, csv, re, sys


def main(src, dst):
    rows = [r for r in csv.reader(l for l in open(src) if l.startswith('"'))]
    hdr = rows[0]
    name_i, val_i = hdr.index("Kernel Name"), hdr.index("Metric Value")
    acc = collections.OrderedDict()
    for r in rows[1:]:
        if len(r) <= val_i:
            continue
        name = re.sub(r"\(.*", "", r[name_i]).replace("void ", "").replace("pg::", "")
        t = float(r[val_i].replace(",", "")) / 1000.0   # ns -> us
        a = acc.setdefault(name, [0, 0.0])
        a[0] += 1
        a[1] += t
    total = sum(a[1] for a in acc.values())
    with open(dst, "w") as f:
        f.write("# %s: %d launches, %.1f us listed (ncu: serialised, cold cache -- compare SHARES, not absolutes)\n" % (src, sum(a[0] for a in acc.values()), total))
        f.write("%-60s %8s %12s %10s %7s\n" % ("kernel", "launches", "total us", "mean us", "share"))
        for name, (n, t) in sorted(acc.items(), key=lambda kv: -kv[1][1]):
            f.write("%-60s %8d %12.1f %10.2f %6.1f%%\n" % (name[:60], n, t, t / n, 100.0 * t / total))
    print(open(dst).read())


if __name__ == "__main__":
    main(sys.argv[1], sys.argv[2])
