"""Row f3 on the device (mbgc_b200/csrc/rcmatch.cu + rcmatch.cpp through mbgcg_rc_match_sequence) against the oracle
restatement of SimpleSequenceMatcher::rcMatchSequence, which tests/test_rcmatch_oracle.py pins against the reference's own
classes.  Plus, where oracle/_ref travels, the reference's own inverse (restoreRCMatchedSequence) on the device's output."""
import numpy as np
import pytest

import mbgc_b200
from rcmatch_cases import RC_CASES, o_rc_match, oracle_rc, r_rc_restore, rc_sequence, ref_copmem, revcomp

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def g():
    m = mbgc_b200.Matcher(1 << 22, 32, 16, 1, 16, device=0)
    yield m
    m.close()


@pytest.mark.parametrize("n,seed", RC_CASES + [(6_000_000, 21)])
@pytest.mark.parametrize("target_len", [55, 32, 120])
def test_rc_match_equals_oracle(g, n, seed, target_len):
    if n > 2_000_000 and target_len != 55:
        pytest.skip("the large case runs at the default length only")
    seq = rc_sequence(n, seed)
    want = o_rc_match(oracle_rc(), seq, target_len)
    s, off, ln, tm = g.rcMatchSequence(seq, target_len)
    assert np.array_equal(off, want[1]), "rcMapOff differs"
    assert np.array_equal(ln, want[2]), "rcMapLen differs"
    assert np.array_equal(s, want[0]), "the rewritten sequence differs"
    ref = ref_copmem()
    if ref is not None:
        assert np.array_equal(r_rc_restore(ref, s, off, ln, n), seq)


def test_long_repeats_are_finished_on_the_host(g):
    """matches far longer than the device's extension cap (4 KiB per side), a low-complexity stretch that over-fills one
    hash bucket, and a match that touches byte 0"""
    rng = np.random.default_rng(77)
    n = 900_000
    seq = np.frombuffer(b"ACGT", dtype=np.uint8)[rng.integers(0, 4, n)].copy()
    seq[500_000:700_000] = revcomp(seq[100_000:300_000])         # a 200 kb reverse-complement repeat
    seq[800_000:830_000] = ord("A")                               # 30 kb of one letter: one bucket with thousands of entries
    seq[n - 3000:] = revcomp(seq[:3000])                          # reaches both ends of the sequence
    want = o_rc_match(oracle_rc(), seq, 55)
    s, off, ln, _ = g.rcMatchSequence(seq, 55)
    assert np.array_equal(off, want[1]) and np.array_equal(ln, want[2]) and np.array_equal(s, want[0])
    assert s.size < n - 190_000


def test_short_and_empty(g):
    for n in (0, 10, 54):
        seq = rc_sequence(n, 3 + n)
        s, off, ln, _ = g.rcMatchSequence(seq, 55)
        assert np.array_equal(s, seq) and off.size == 0 and ln.size == 0
    with pytest.raises(mbgc_b200.MatcherError):
        g.rcMatchSequence(rc_sequence(1000, 1), 20)               # "Minimal matching length too short"
