"""Row f3: the oracle restatement of SimpleSequenceMatcher::rcMatchSequence (oracle/mbgc_oracle.c: mbo_rc_match_sequence)
against the REFERENCE's own classes compiled from where they lie (oracle/_ref/libref_copmem.so, single-threaded index
build) and against golden digests generated from them.  CPU only."""
import hashlib
import json
import os

import numpy as np
import pytest

from rcmatch_cases import RC_CASES, o_rc_match, oracle_rc, r_rc_match, r_rc_restore, rc_sequence, ref_copmem
from util import GOLDEN

REF = ref_copmem()
needs_ref = pytest.mark.skipif(REF is None, reason="oracle/_ref/libref_copmem.so not built (no /root/reference here)")


def sha(*arrs):
    h = hashlib.sha1()
    for a in arrs:
        h.update(bytes(a))
        h.update(b"|")
    return h.hexdigest()[:16]


@needs_ref
@pytest.mark.parametrize("n,seed", RC_CASES)
@pytest.mark.parametrize("target_len", [55, 32, 120])
def test_rc_match_equals_reference(n, seed, target_len):
    seq = rc_sequence(n, seed)
    got = o_rc_match(oracle_rc(), seq, target_len)
    want = r_rc_match(REF, seq, target_len)
    assert got is not None
    assert np.array_equal(got[1], want[1]), "rcMapOff differs"
    assert np.array_equal(got[2], want[2]), "rcMapLen differs"
    assert np.array_equal(got[0], want[0]), "the rewritten sequence differs"
    if n >= 3000 and target_len == 55:
        assert got[1].size > 0, "the case is meant to contain reverse-complement matches"
    back = r_rc_restore(REF, got[0], got[1], got[2], n)    # the reference's own inverse restores the input
    assert np.array_equal(back, seq)


def test_short_sequence_is_untouched():
    seq = rc_sequence(40, 9)
    s, off, ln, raw = o_rc_match(oracle_rc(), seq, 55)
    assert np.array_equal(s, seq) and off.size == 0 and ln.size == 0


def test_golden_digests():
    g = json.load(open(os.path.join(GOLDEN, "golden.json"))).get("rcmatch")
    if g is None:
        pytest.skip("golden.json has no rcmatch section yet")
    L = oracle_rc()
    for key, want in g.items():
        n, seed, tl = (int(x) for x in key.split(":"))
        got = o_rc_match(L, rc_sequence(n, seed), tl)
        assert sha(got[0], got[1], got[2]) == want
