"""Summarise an `ncu --page source --csv` dump: instruction mix and the hottest SASS lines by stall samples.
usage: python profiles/ncu_src_summary.py probe_src.csv [top_n]"""
import collections
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
top = int(sys.argv[2]) if len(sys.argv) > 2 else 25
hdr_i = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
hdr = rows[hdr_i]
col = {h: i for i, h in enumerate(hdr)}
body = []
for r in rows[hdr_i + 1:]:
    if r and r[0] == "Kernel Name":
        break  # first launch only
    if len(r) == len(hdr):
        body.append(r)
num = lambda r, c: float(r[col[c]] or 0)
tot_inst = sum(num(r, "Instructions Executed") for r in body)
tot_samp = sum(num(r, "# Samples") for r in body)
print(f"SASS lines {len(body)}  warp-instructions executed {tot_inst:.3e}  stall samples {tot_samp:.0f}")
mix = collections.Counter()
smp = collections.Counter()
for r in body:
    op = r[col["Source"]].split()[0] if r[col["Source"]].split() else "?"
    if op.startswith("@"):
        op = r[col["Source"]].split()[1]
    op = op.split(".")[0]
    mix[op] += num(r, "Instructions Executed")
    smp[op] += num(r, "# Samples")
print("opcode mix (share of executed warp instructions | share of stall samples):")
for op, v in mix.most_common(18):
    print(f"  {op:10s} {v / tot_inst * 100:6.2f}%  | {smp[op] / max(tot_samp, 1) * 100:6.2f}%")
stalls = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
agg = {s: sum(num(r, s) for r in body) for s in stalls}
print("stall reasons (all samples):", ", ".join(f"{k[6:]} {v / max(tot_samp, 1) * 100:.1f}%" for k, v in
                                              sorted(agg.items(), key=lambda x: -x[1])[:8]))
print(f"hottest {top} lines by samples:")
for r in sorted(body, key=lambda r: -num(r, "# Samples"))[:top]:
    st = sorted(((num(r, s), s[6:]) for s in stalls), reverse=True)[:2]
    print(f"  {num(r, '# Samples') / max(tot_samp, 1) * 100:5.2f}%  exec {num(r, 'Instructions Executed'):.2e}  "
          f"{r[col['Source']][:70]:70s} {st[0][1]}:{st[0][0]:.0f} {st[1][1]}:{st[1][0]:.0f}")
