import sys, time
sys.path.insert(0, "/root/repo")
import numpy as np, torch
import mbgc_b200
from mbgc_b200.synth import synth_base, synth_mutant, mgmp_ref_limit
base = synth_base(5_000_000, 1234)
texts = [synth_mutant(base, 1234 * 1000003 + i, 1e-4, 1e-5) for i in range(1, 17)]
m = mbgc_b200.Matcher(mgmp_ref_limit(33, base.size), 32, 16, 1, 16)
m.disableSlidingWindow(); m.loadRef(base, True, True, 0)
pinned = [torch.from_numpy(t).pin_memory().numpy() for t in texts]
for label, src in (("pageable", texts), ("pinned", pinned), ("pageable", texts)):
    for k in range(4):
        t0 = time.perf_counter(); m.matchBatch(src, 32, fetch=False); dt = time.perf_counter() - t0
        print(label, k, round(dt * 1e3, 2), "ms", {a: round(b, 3) for a, b in m.timing().items() if b > 0.01}, flush=True)
