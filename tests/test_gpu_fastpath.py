"""GPU parity of the fast path for near-identical texts (kernels.cu: self-probe memo, k_anchor, k_clean).

The fast path may only ever declare a 512-position chunk "clean" when the full probe would return exactly the
seeds it synthesises, so every result must stay bit-identical to the CPU oracle -- with the fast path forced on
(MBGC_B200_FASTPATH=1), forced off (=0) and under the cost model (unset).  The reference semantics checked are
those of matching/SlidingWindowSparseEMMatcher.cpp:181-321 via the oracle."""
import numpy as np
import pytest

from scenarios import CudaFace, all_scenarios, replay
from util import OracleMatcher, mgmp_ref_limit, synth_collection

pytestmark = pytest.mark.gpu


@pytest.fixture(params=["1", "0"], ids=["fast-forced", "fast-off"])
def fastpath(request, monkeypatch):
    monkeypatch.setenv("MBGC_B200_FASTPATH", request.param)
    return request.param


@pytest.mark.parametrize("skip_margin", [16, 24])
@pytest.mark.parametrize("sc", all_scenarios(), ids=lambda s: s["name"])
def test_scenarios_forced(sc, skip_margin, fastpath):
    o = OracleMatcher(sc["limit"], 32, 16, 1, skip_margin)
    g = CudaFace(sc["limit"], 32, 16, 1, skip_margin)
    try:
        want, sw = replay(o, sc)
        got, sg = replay(g, sc)
        assert sg == sw
        assert len(got) == len(want)
        for k, (a, b) in enumerate(zip(got, want)):
            assert np.array_equal(a, b), f"op {k}: cuda {len(a)} matches, oracle {len(b)}"
    finally:
        g.close()
        o.close()


def test_fast_path_engages_and_matches(monkeypatch):
    """near tier: almost every chunk must leave through the fast path, and the lists must equal the oracle's"""
    monkeypatch.delenv("MBGC_B200_FASTPATH", raising=False)
    gs = synth_collection(9, 1_000_000, 99, 1e-4, 1e-5)
    limit = mgmp_ref_limit(9, gs[0].size)
    o = OracleMatcher(limit)
    g = CudaFace(limit)
    try:
        for m in (o, g):
            m.disable_sliding_window()
            m.load_ref(gs[0], rc=True)
        got = g.match_batch(gs[1:])
        c = g.m.counters()
        assert c["chunks"] > 0 and c["lanes_full_probe"] < 0.05 * 32 * c["chunks"], c
        for a, t in zip(got, gs[1:]):
            assert np.array_equal(a, o.match(t))
        # the memo survives separators (no index change) and is rebuilt after a reference extension
        for m in (o, g):
            m.load_separator(0)
        assert np.array_equal(g.match(gs[1]), o.match(gs[1]))
        for m in (o, g):
            m.load_ref(gs[2], rc=True)
            m.load_separator(0)
        got = g.match_batch(gs[3:])
        for a, t in zip(got, gs[3:]):
            assert np.array_equal(a, o.match(t))
    finally:
        g.close()
        o.close()


def test_divergent_batch_stays_on_full_probe(monkeypatch):
    """div tier: the fast path may stay on (lanes between SNPs are clean) or switch off; lists must not care"""
    monkeypatch.delenv("MBGC_B200_FASTPATH", raising=False)
    gs = synth_collection(5, 400_000, 7, 1e-2, 1e-4)
    limit = mgmp_ref_limit(5, gs[0].size)
    o = OracleMatcher(limit)
    g = CudaFace(limit)
    try:
        for m in (o, g):
            m.disable_sliding_window()
            m.load_ref(gs[0], rc=True)
        for rnd in range(2):
            got = g.match_batch(gs[1:])
            for a, t in zip(got, gs[1:]):
                assert np.array_equal(a, o.match(t))
        c = g.m.counters()
        assert 0 < c["lanes_full_probe"] <= 32 * c["chunks"], c
    finally:
        g.close()
        o.close()


@pytest.mark.parametrize("lock_mode", ["acquire", None])
def test_sliding_window_classes(lock_mode, fastpath):
    """window classes (ref .cpp:212-222) with worker locks: clean chunks must not straddle a class boundary"""
    gs = synth_collection(6, 150_000, 5, 3e-4, 3e-5)
    limit = 1_000_003
    o = OracleMatcher(limit)
    g = CudaFace(limit)
    try:
        for m in (o, g):
            m.set_sliding_window_size(16)
            m.load_ref(gs[0], rc=True)
        for t in gs[1:]:
            res = []
            for m in (g, o):
                if lock_mode:
                    lock = m.acquire_lock()
                    res.append(m.match(t, 32, lock))
                    m.release_lock(lock)
                else:
                    res.append(m.match(t))
                m.load_ref(t, rc=False)
                m.load_separator(0)
            assert np.array_equal(res[0], res[1])
    finally:
        g.close()
        o.close()


@pytest.mark.parametrize("pass_bytes", ["300000", "2500000"])
def test_many_passes_equal_one(monkeypatch, pass_bytes):
    """a batch cut into many passes (MBGC_B200_MAX_PASS_BYTES) gives the lists of the single pass: per-pass
    buffers, the pinned result buffer that grows across passes and the guessed match copy are all exercised"""
    monkeypatch.delenv("MBGC_B200_FASTPATH", raising=False)
    gs = synth_collection(10, 600_000, 17, 1e-3, 1e-4)
    ts = gs[1:] + [gs[1][:10], np.zeros(0, np.uint8), gs[2][:100_000]]
    limit = mgmp_ref_limit(10, gs[0].size)
    g = CudaFace(limit)
    try:
        g.disable_sliding_window()
        g.load_ref(gs[0], rc=True)
        one = g.match_batch(ts)
        assert g.m.counters()["passes"] == 1
        monkeypatch.setenv("MBGC_B200_MAX_PASS_BYTES", pass_bytes)
        many = g.match_batch(ts)
        assert g.m.counters()["passes"] > 2
        again = g.match_batch(ts)
        for a, b, c in zip(one, many, again):
            assert np.array_equal(a, b) and np.array_equal(a, c)
    finally:
        g.close()


@pytest.mark.parametrize("L,sm", [(20, 16), (24, 24), (32, 16), (40, 0), (48, 16), (56, 16), (64, 0), (120, 16)])
def test_fast_path_other_kmer_lengths(monkeypatch, L, sm):
    """k1 = 16 with K = 16 .. 56 (K-1 tail bytes of 15 .. 55, one or two tail words per lane, 2 .. 5 pieces per
    lane): near-identical texts with repeats, fast path forced on, against the oracle"""
    monkeypatch.setenv("MBGC_B200_FASTPATH", "1")
    from scenarios import repeat_rich_near
    sc = repeat_rich_near(seed=43, length=150_000, n=3)
    o = OracleMatcher(sc["limit"], L, 16, 1, sm)
    g = CudaFace(sc["limit"], L, 16, 1, sm)
    try:
        want, sw = replay(o, sc, min_len=L)
        got, sg = replay(g, sc, min_len=L)
        assert sg == sw
        for k, (a, b) in enumerate(zip(got, want)):
            assert np.array_equal(a, b), f"op {k}: cuda {len(a)} matches, oracle {len(b)}"
        c = g.m.counters()
        assert c["lanes_full_probe"] < 32 * c["chunks"], c  # the fast path did run
    finally:
        g.close()
        o.close()
