"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list into per-kernel totals and shares.
usage: python profiles/summarize_launches.py gpurun_out/launches.csv > profiles/launches_rNN.md"""
import collections
import csv
import re
import sys


def main(path):
    rows = [r for r in csv.reader(open(path)) if len(r) > 14 and r[0] != "ID"]
    agg = collections.OrderedDict()
    for r in rows:
        name = re.sub(r"^void\s+", "", r[4])
        name = re.sub(r"<unnamed>::", "", name)
        name = re.sub(r"\(.*$", "", name)
        ns = float(r[14])
        a = agg.setdefault(name, [0, 0.0])
        a[0] += 1; a[1] += ns
    total = sum(v[1] for v in agg.values())
    print(f"# ncu launch list summary: {path}\n")
    print(f"launches: {len(rows)}, summed device time: {total/1e6:.3f} ms (cold-cache, serialised: compare shares, not absolutes)\n")
    print("| kernel | launches | total ms | mean us | share |")
    print("|---|---:|---:|---:|---:|")
    for name, (cnt, ns) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print(f"| `{name}` | {cnt} | {ns/1e6:.3f} | {ns/cnt/1e3:.1f} | {100*ns/total:.1f}% |")


if __name__ == "__main__":
    main(sys.argv[1])
