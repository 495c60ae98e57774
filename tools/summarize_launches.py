"""Summarise an `ncu --csv --metrics gpu__time_duration.sum,...` launch list per kernel.
Usage: python tools/summarize_launches.py launches.csv steps > profiles/x.md"""
import collections, csv, sys

path, steps = sys.argv[1], int(sys.argv[2]) if len(sys.argv) > 2 else 1
rows = list(csv.reader(open(path)))
h = next(i for i, r in enumerate(rows) if r and r[0] == "ID")
hdr, data = rows[h], rows[h + 1:]
ki, mi, vi, ii = hdr.index("Kernel Name"), hdr.index("Metric Name"), hdr.index("Metric Value"), hdr.index("ID")
per = collections.OrderedDict()
for r in data:
    if len(r) > vi:
        per.setdefault((r[ii], r[ki]), {})[r[mi]] = float(r[vi].replace(",", ""))
agg = collections.OrderedDict()
for (_, name), m in per.items():
    a = agg.setdefault(name.split("(")[0].replace("void ", ""), [0, 0.0, 0.0, 0.0])
    a[0] += 1; a[1] += m.get("gpu__time_duration.sum", 0); a[2] += m.get("dram__bytes_read.sum", 0); a[3] += m.get("dram__bytes_write.sum", 0)
tot = sum(a[1] for a in agg.values())
print(f"| kernel | launches/step | ms/step | share | DRAM read MB/step | DRAM write MB/step | DRAM GB/s |")
print("|---|---:|---:|---:|---:|---:|---:|")
for name, a in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    ms, rd, wr = a[1] / 1e6 / steps, a[2] / 1e6 / steps, a[3] / 1e6 / steps
    print(f"| {name} | {a[0] / steps:g} | {ms:.3f} | {a[1] / tot * 100:.1f}% | {rd:.1f} | {wr:.1f} | {(rd + wr) / ms if ms > 0 else 0:.0f} |")
print(f"| **total** | {sum(a[0] for a in agg.values()) / steps:g} | {tot / 1e6 / steps:.3f} | 100% | | | |")
