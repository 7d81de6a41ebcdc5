#!/usr/bin/env python
"""matchTexts of one 5 Mbp near-identical contig against a GROWN reference (G0 + two divergent genomes appended, each
with its reverse complement): the regime where seeds of several diagonals interleave (profiles/README.md).  Prints the
stage times; run it under `ncu --metrics gpu__time_duration.sum` for the per-kernel list."""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import mbgc_b200  # noqa: E402
from mbgc_b200.synth import mgmp_ref_limit, synth_base, synth_mutant  # noqa: E402

base = synth_base(5_000_000, 1234)
m = mbgc_b200.Matcher(mgmp_ref_limit(33, base.size), 32, 16, 1, 16)
m.disableSlidingWindow()
t0 = time.perf_counter()
m.loadRef(base, True, True, 0)
print("loadRef G0 + RC", round((time.perf_counter() - t0) * 1e3, 2), "ms", flush=True)
for i in (3, 10):
    g = synth_mutant(base, 1234 * 1000003 + i, 1e-2, 1e-4)
    t0 = time.perf_counter()
    m.loadRef(g, True, True, 0)
    print("loadRef 5 Mbp + RC", round((time.perf_counter() - t0) * 1e3, 2), "ms", flush=True)
t = synth_mutant(base, 1234 * 1000003 + 1, 1e-4, 1e-5)
for k in range(int(sys.argv[1]) if len(sys.argv) > 1 else 4):
    t0 = time.perf_counter()
    r = m.matchTexts(t, 32)
    print(k, round((time.perf_counter() - t0) * 1e3, 2), "ms", len(r), "matches",
          {a: round(b, 3) for a, b in m.timing().items() if b > 0.01}, m.counters(), flush=True)
