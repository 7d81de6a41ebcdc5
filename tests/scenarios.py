"""Seeded matcher scenarios shared by the CPU (oracle vs reference) and GPU (CUDA vs oracle) parity tests.

A scenario is a list of operations replayed on any matcher face with the methods of tests/util._MatcherBase:
  ("load", text, rc)            loadRef
  ("sep",)                      loadSeparator
  ("match", text, lock|None)    matchTexts   (lock = "acquire" takes a worker lock first and releases it after)
  ("batch", [texts])            n matchTexts against the same frozen state (GPU: one mbgcg_match_batch call)
"""
import numpy as np

from util import synth_base, synth_mutant

_ACGT = np.frombuffer(b"ACGT", dtype=np.uint8)


def _rng(seed):
    return np.random.default_rng(seed)


def near_collection(seed=11, n=4, length=200_000, p_snp=1e-3, p_indel=1e-4):
    base = synth_base(length, seed)
    ops = [("load", base, True)]
    for i in range(1, n):
        g = synth_mutant(base, seed * 7919 + i, p_snp, p_indel)
        ops.append(("match", g, None))
    return dict(limit=1 << 23, ops=ops, name=f"near{seed}")


def growing_collection(seed=5, n=5, length=120_000, p_snp=1e-2, p_indel=1e-3):
    """every contig (and its RC) is appended after matching: the reference keeps growing"""
    base = synth_base(length, seed)
    ops = [("load", base, True)]
    for i in range(1, n):
        g = synth_mutant(base, seed * 7919 + i, p_snp, p_indel)
        ops += [("match", g, None), ("load", g, True), ("sep",)]
    return dict(limit=1 << 22, ops=ops, name=f"grow{seed}")


def wrapping_collection(seed=9, n=9, length=90_000, limit=400_001, sliding=False):
    """reference buffer far smaller than the loaded data: circular-buffer wrap-arounds, post-wrap sampling
    phase (samplingPos reset to 1), stale entries, window classes"""
    base = synth_base(length, seed)
    ops = [("load", base, True)]
    for i in range(1, n):
        g = synth_mutant(base, seed * 7919 + i, 5e-3, 5e-4)
        ops += [("match", g, "acquire" if sliding else None), ("load", g, i % 2 == 0), ("sep",)]
    return dict(limit=limit, ops=ops, name=f"wrap{seed}{'sw' if sliding else ''}", sliding=sliding)


def pathological(seed=3):
    """low complexity, N runs, lower case, short texts, repeats (swallow / overlap logic)"""
    r = _rng(seed)
    unit = _ACGT[r.integers(0, 4, 37)]
    tandem = np.tile(unit, 400)
    rnd = _ACGT[r.integers(0, 4, 30_000)]
    polyA = np.full(5_000, ord("A"), dtype=np.uint8)
    nrun = np.full(3_000, ord("N"), dtype=np.uint8)
    lower = np.frombuffer(bytes(rnd[:4000]).lower(), dtype=np.uint8)
    ref = np.concatenate([rnd, polyA, rnd[:10_000], nrun, tandem, lower, rnd[5_000:25_000]])
    t1 = np.concatenate([rnd[100:20_000], nrun[:1500], polyA[:2000], tandem[:5000], rnd[:3000], lower[:2000]])
    t2 = np.concatenate([tandem[11:9000], polyA, rnd[20_000:], nrun, rnd[:50]])
    dup = np.concatenate([rnd[:8000], rnd[:8000], rnd[4000:12000]])
    ops = [("load", ref, True), ("match", t1, None), ("match", t2, None), ("match", dup, None),
           ("load", t1, True), ("match", t2, None), ("match", dup, None),
           ("match", rnd[:27], None), ("match", rnd[:28], None), ("match", rnd[:31], None), ("match", rnd[:32], None),
           ("match", rnd[:33], None), ("match", np.zeros(0, np.uint8), None),
           ("batch", [t1, rnd[:10], t2, np.zeros(0, np.uint8), dup, rnd[:64]])]
    return dict(limit=1 << 22, ops=ops, name=f"patho{seed}")


def many_small_contigs(seed=21, n=300):
    r = _rng(seed)
    base = synth_base(150_000, seed)
    texts = []
    for i in range(n):
        ln = int(r.integers(20, 3000))
        st = int(r.integers(0, base.size - ln))
        t = base[st:st + ln].copy()
        k = int(r.integers(0, 4))
        if k and ln > 40:
            t[r.integers(0, ln, k)] = ord("A")
        texts.append(t)
    return dict(limit=1 << 22, ops=[("load", base, True), ("batch", texts)], name=f"small{seed}")


def repeat_rich_near(seed=41, length=400_000, n=4):
    """near-identical targets of a reference full of repeats: duplicated elements (a later copy overwrites the
    table slot of an earlier one), tandem repeats, homopolymers, N runs, a lower-case stretch.  Exercises the
    fast path's memo (which chunks may be declared clean) against repeats on and off the sampling grid."""
    r = _rng(seed)
    base = synth_base(length, seed).copy()
    at = lambda x: x * length // 400_000              # positions are given for the default length
    elem = base[at(50_000):at(50_000) + 6_000].copy()
    for p in (120_003, 200_010, 300_007):       # copies at different phases of the 16-grid
        base[at(p):at(p) + elem.size] = elem
    unit = _ACGT[r.integers(0, 4, 48)]           # tandem repeat, period 48 = 3 x 16
    base[at(250_000):at(250_000) + 48 * 60] = np.tile(unit, 60)
    base[at(20_000):at(20_000) + 1_500] = ord("A")
    base[at(330_000):at(330_000) + 1_000] = ord("N")
    lo = at(350_000)
    base[lo:lo + 2_000] = np.frombuffer(bytes(base[lo:lo + 2_000]).lower(), dtype=np.uint8)
    ops = [("load", base, True)]
    texts = [synth_mutant(base, seed * 7919 + i, 2e-4, 2e-5) for i in range(1, n)]
    texts.append(base.copy())                                    # identical copy
    texts.append(np.concatenate([base[77:at(150_000)], base[10:at(90_000)]]))  # rearranged, off-grid
    ops.append(("batch", texts))
    ops += [("sep",), ("match", texts[0], None), ("load", texts[1], True), ("match", texts[2], None), ("sep",),
            ("match", texts[0], None), ("batch", texts[:3])]
    return dict(limit=1 << 23, ops=ops, name=f"repeats{seed}")


def all_scenarios():
    return [near_collection(), near_collection(seed=12, n=3, length=300_000, p_snp=1e-4, p_indel=1e-5),
            growing_collection(), wrapping_collection(), wrapping_collection(seed=10, sliding=True, limit=700_003),
            pathological(), many_small_contigs(), repeat_rich_near()]


def replay(m, sc, batch=True, min_len=32):
    """Runs a scenario on matcher face `m` (methods of util._MatcherBase or the adapter below); returns the list
    of match arrays in operation order plus the final (refLength, loadingPosition, loadedRefLength)."""
    if sc.get("sliding"):
        m.set_sliding_window_size(16)
    else:
        m.disable_sliding_window()
    out = []
    for op in sc["ops"]:
        if op[0] == "load":
            m.load_ref(op[1], rc=op[2])
        elif op[0] == "sep":
            m.load_separator(0)
        elif op[0] == "match":
            if op[2] == "acquire":
                lock = m.acquire_lock()
                out.append(m.match(op[1], min_len, lock))
                m.release_lock(lock)
            else:
                out.append(m.match(op[1], min_len))
        elif op[0] == "batch":
            if batch and hasattr(m, "match_batch"):
                out.extend(m.match_batch(op[1], min_len))
            else:
                out.extend(m.match(t, min_len) for t in op[1])
    return out, (m.ref_length, m.loading_position, m.loaded_ref_length)


class CudaFace:
    """adapter: mbgc_b200.Matcher with the snake_case face the scenarios use"""

    def __init__(self, limit, L=32, k1=16, k2=1, skip_margin=16):
        import mbgc_b200
        self.m = mbgc_b200.Matcher(limit, L, k1, k2, skip_margin)

    def close(self):
        self.m.close()

    def disable_sliding_window(self):
        self.m.disableSlidingWindow()

    def disable_circular_buffer(self):
        self.m.disableCircularBuffer()

    def set_sliding_window_size(self, f):
        self.m.setSlidingWindowSize(f)

    def load_ref(self, text, rc=False, add_sep=True, sep=0):
        self.m.loadRef(text, rc, add_sep, sep)

    def load_separator(self, sep=0):
        self.m.loadSeparator(sep)

    def match(self, text, min_len=32, lock=2**64 - 1):
        return self.m.matchTexts(text, min_len, lock)

    def match_batch(self, texts, min_len=32, locks=None):
        return self.m.matchBatch(texts, min_len, locks)

    def acquire_lock(self):
        return self.m.acquireWorkerMatchingLockPos()

    def release_lock(self, v):
        self.m.releaseWorkerMatchingLockPos(v)

    ref_length = property(lambda s: s.m.getRefLength())
    loading_position = property(lambda s: s.m.getLoadingPosition())
    loaded_ref_length = property(lambda s: s.m.getLoadedRefLength())
    max_ref_length = property(lambda s: s.m.getMaxRefLength())

    def table(self):
        return self.m.table()

    def ref(self, n=None):
        return self.m.getRef(n)
