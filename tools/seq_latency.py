#!/usr/bin/env python
"""Per-call latency of the sequential (one contig per matchTexts call) use of the matcher, as the drop-in drives it."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import mbgc_b200
from mbgc_b200.synth import synth_base, synth_mutant, mgmp_ref_limit

base = synth_base(5_000_000, 1234)
ts = [synth_mutant(base, 1234 * 1000003 + i, 1e-4, 1e-5) for i in range(1, 13)]
m = mbgc_b200.Matcher(mgmp_ref_limit(101, base.size), 32, 16, 1, 16)
m.disableSlidingWindow()
t0 = time.perf_counter(); m.loadRef(base, True, True, 0); print("loadRef G0+RC %.1f ms" % ((time.perf_counter() - t0) * 1e3))
for i, t in enumerate(ts):
    t0 = time.perf_counter()
    r = m.matchTexts(t, 32)
    dt = (time.perf_counter() - t0) * 1e3
    tm = m.timing()
    print(f"call {i}: {dt:7.2f} ms wall, {len(r)} matches, stages " + " ".join(f"{k}={v:.3f}" for k, v in tm.items() if v > 0.0005), m.counters()["lanes_full_probe"])
    t0 = time.perf_counter(); m.loadRef(np.zeros(0, np.uint8), False, True, 0); m.loadSeparator(0)
    print("   empty loadRef + separator %.3f ms" % ((time.perf_counter() - t0) * 1e3))
