#!/usr/bin/env python
"""Per-kernel SASS opcode histogram of libmbgc_b200.so (cuobjdump -sass): what proves which hardware paths the kernels
use (UBLKCP / SYNCS = TMA bulk copies behind mbarriers, MATCH = __match_any_sync, ATOM / RED, REDUX, SHFL, VOTE ...).

    python profiles/sass_histogram.py > profiles/r02/sass_opcode_histogram.txt
"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "mbgc_b200", "libmbgc_b200.so")
KEY = ["LDG", "STG", "LDS", "STS", "ATOM", "ATOMG", "ATOMS", "RED", "REDUX", "MATCH", "SHFL", "VOTE", "BAR", "UBLKCP", "SYNCS",
       "UTMALDG", "LOP3", "SHF", "IMAD", "PRMT", "VIADD", "BRA"]


def main():
    out = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True).stdout
    kernels = collections.OrderedDict()
    cur = None
    arch = set()
    for line in out.splitlines():
        m = re.search(r"Function : (\S+)", line)
        if m:
            name = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip()
            cur = kernels.setdefault(re.sub(r"\(.*", "", name), collections.Counter())
            continue
        m = re.search(r"arch = (sm_\w+)", line)
        if m:
            arch.add(m.group(1))
        m = re.match(r"\s+/\*[0-9a-f]{4}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_]+)", line)
        if m and cur is not None:
            op = m.group(1)
            cur[op.split(".")[0]] += 1
            cur["_total"] += 1
    print(f"# {os.path.relpath(LIB, ROOT)}: {len(kernels)} kernels, arch {sorted(arch)}")
    print(f"{'kernel':58s} {'instr':>6s} " + " ".join(f"{k:>6s}" for k in KEY))
    for name, c in kernels.items():
        print(f"{name[:58]:58s} {c['_total']:6d} " + " ".join(f"{c.get(k, 0):6d}" for k in KEY))
    tot = collections.Counter()
    for c in kernels.values():
        tot.update(c)
    print(f"{'all kernels':58s} {tot['_total']:6d} " + " ".join(f"{tot.get(k, 0):6d}" for k in KEY))


if __name__ == "__main__":
    main()
