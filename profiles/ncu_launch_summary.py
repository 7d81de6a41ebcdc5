"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list: per-kernel count, total time, share.
usage: python profiles/ncu_launch_summary.py launches.csv"""
import collections
import csv
import sys

rows = list(csv.reader(l for l in open(sys.argv[1]) if l.startswith('"')))
hdr = rows[0]
ki, vi, ui = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
agg = collections.OrderedDict()
for r in rows[1:]:
    k = r[ki].split("(")[0].replace("mbgc_b200::", "").replace("void ", "")
    v = float(r[vi].replace(",", ""))
    v = v / 1000 if r[ui] in ("ns", "nsecond") else (v * 1000 if r[ui] in ("ms", "msecond") else v)
    agg.setdefault(k, [0.0, 0])
    agg[k][0] += v
    agg[k][1] += 1
tot = sum(v[0] for v in agg.values())
print(f"{'kernel':50s} {'launches':>8s} {'total us':>10s} {'share':>7s}")
for k, (v, n) in sorted(agg.items(), key=lambda x: -x[1][0]):
    print(f"{k[:50]:50s} {n:8d} {v:10.1f} {v / tot * 100:6.1f}%")
