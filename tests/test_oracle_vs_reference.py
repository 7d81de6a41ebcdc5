"""The CPU oracle against the reference's own matcher class, live (oracle/_ref/libref_matcher.so = the
reference's SlidingWindowSparseEMMatcher.cpp compiled unmodified).  Skipped where oracle/_ref is absent."""
import numpy as np
import pytest

from scenarios import all_scenarios, growing_collection, replay, wrapping_collection
from util import OracleMatcher, RefMatcher, have_ref_driver

pytestmark = pytest.mark.skipif(not have_ref_driver(), reason="oracle/_ref/libref_matcher.so not built")


def _same(sc, L=32, k1=16, sm=16):
    o, r = OracleMatcher(sc["limit"], L, k1, 1, sm), RefMatcher(sc["limit"], L, k1, 1, sm)
    a, sa = replay(o, sc, min_len=L)
    b, sb = replay(r, sc, min_len=L)
    assert sa == sb and o.kmer == r.kmer and o.sampling_pos == r.sampling_pos and o.table_size == r.table_size
    assert len(a) == len(b)
    for x, y in zip(a, b):
        assert np.array_equal(x, y)
    assert np.array_equal(o.table(), r.table())
    n = o.ref_length
    assert np.array_equal(o.ref(n), r.ref(n))
    o.close()
    r.close()


@pytest.mark.parametrize("skip_margin", [16, 24])
@pytest.mark.parametrize("sc", all_scenarios(), ids=lambda s: s["name"])
def test_scenarios(sc, skip_margin):
    _same(sc, sm=skip_margin)


@pytest.mark.parametrize("L,k1,sm", [(32, 7, 16), (32, 8, 24), (32, 12, 16), (32, 6, 0), (32, 32, 16), (24, 16, 24),
                                      (20, 16, 16), (16, 16, 16), (16, 4, 0), (48, 16, 16), (64, 16, 0), (40, 5, 3),
                                      (120, 16, 16)])
def test_parameter_variants(L, k1, sm):
    _same(growing_collection(seed=31, n=4, length=60_000), L, k1, sm)
    _same(wrapping_collection(seed=32, n=6, length=50_000, limit=260_001), L, k1, sm)


def test_locks_protocol():
    """acquire / release order, erased entries, swEnd updates (ref .cpp:361-400)"""
    o, r = OracleMatcher(1_000_003), RefMatcher(1_000_003)
    rng = np.random.default_rng(1)
    text = np.frombuffer(b"ACGT", dtype=np.uint8)[rng.integers(0, 4, 200_000)]
    for m in (o, r):
        m.set_sliding_window_size(16)
    held_o, held_r = [], []
    for step in range(40):
        for m in (o, r):
            m.load_ref(text[step * 4000:(step + 1) * 4000 + 30_000], rc=step % 3 == 0)
        lo, lr = o.acquire_lock(), r.acquire_lock()
        assert lo == lr
        held_o.append(lo)
        held_r.append(lr)
        if step % 2 and held_o:
            k = int(rng.integers(0, len(held_o)))
            o.release_lock(held_o.pop(k))
            r.release_lock(held_r.pop(k))
        assert (o.ref_length, o.loading_position, o.loaded_ref_length) == (r.ref_length, r.loading_position, r.loaded_ref_length)
    assert np.array_equal(o.table(), r.table())
