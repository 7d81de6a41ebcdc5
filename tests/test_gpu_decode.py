"""GPU parity of row f4 (the decoder's match-copy path): mbgcg_decode_targets against the oracle's restatement of
MBGC_Decoder::decodeTarget & co (oracle/mbgc_oracle_dec.c, pinned against the reference's own class in
tests/test_decode_oracle.py) -- decoded contigs, unmatched counts, reference buffer and cursor, bit-exact; and the round trip
encode (oracle of MBGC_Encoder) -> device decode = the bytes that were encoded."""
import os

import numpy as np
import pytest

from decode_cases import STREAMS, encode_collection, oracle_decode, synth_files
from util import FIXTURES, PackParams as OraclePackParams, have_fixtures, read_fasta

pytestmark = pytest.mark.gpu


def _pp(enable=1, min_len=256, depth=64, gdepth=2, lazy=1, frugal=1, bit40=0, excl=1, init=125, thr=500, pen=50, bonus=50):
    return OraclePackParams(enable, min_len, depth, gdepth, lazy, frugal, bit40, excl, init, thr, pen, bonus, 32, 128)


def _dev_params(pp):
    from mbgc_b200.matcher import PackParams
    return PackParams(pp.enableExtensionsWithMismatches, pp.gapDepthOffsetEncoding, pp.gapDepthMismatchesEncoding,
                      pp.lazyDecompressionSupport, pp.frugal64bitLenEncoding, pp.enable40bitReference, pp.mismatchesWithExclusion,
                      pp.mmsMismatchesInitialScore, pp.mmsMismatchesScoreThreshold, pp.mmsMismatchPenalty, pp.mmsMatchBonus, 0,
                      pp.gapBreakingMatchMinLength)


_matcher = None


def _gpu():
    global _matcher
    if _matcher is None:
        import mbgc_b200
        _matcher = mbgc_b200.Matcher(1 << 22, 32, 16, 1, 16)
    return _matcher


def _device_decode(enc, locks=None, cir=True):
    g = _gpu()
    tg = enc.targets
    return g.decodeTargets(_dev_params(enc.pp), [enc.streams[n] for n in STREAMS], enc.ref_total, [c for _, c in tg], enc.factors,
                           g0IsTarget=enc.g0_is_target, g0SeqCount=enc.seq_counts[0], rcInReference=enc.rc_in_ref,
                           contigsIndividuallyReversed=cir, lockPos=locks)


def _check(enc, locks=None, cir=True, round_trip=True):
    got = _device_decode(enc, locks, cir)
    res, g0, ref, ref_pos, cur = oracle_decode(enc, locks, cir)
    want = ([] if g0 is None else list(g0)) + [c for parts, _ in res for c in parts]
    want_unm = ([0] * (0 if g0 is None else len(g0))) + [u for _, unm in res for u in unm]
    assert len(got["contigs"]) == len(want)
    for i, (a, b) in enumerate(zip(got["contigs"], want)):
        assert a.size == b.size, f"contig {i}: length {a.size} vs {b.size}"
        if not np.array_equal(a, b):
            d = np.nonzero(a != b)[0]
            raise AssertionError(f"contig {i}: {d.size} bytes differ, first at {d[0]} of {a.size}: cuda {bytes(a[d[0]:d[0] + 12])} oracle {bytes(b[d[0]:d[0] + 12])}")
    assert [int(x) for x in got["unmatched"]] == [int(x) for x in want_unm]
    assert got["refPos"] == ref_pos
    dref = _gpu().decodeFetchRef(0, enc.ref_total)
    assert np.array_equal(dref, ref), "the decoder's reference buffer differs from the oracle's"
    if round_trip:
        flat = [c for f, _ in enc.targets for c in enc.contigs[f]]
        if not enc.g0_is_target:
            flat = list(enc.contigs[0]) + flat
        for i, (a, b) in enumerate(zip(got["contigs"], flat)):
            assert np.array_equal(a, b), f"contig {i}: the decoded bytes are not the encoded ones"
    return got


VARIANTS = [dict(), dict(frugal=0), dict(lazy=0), dict(enable=0), dict(depth=0), dict(depth=3), dict(depth=127), dict(gdepth=0),
            dict(gdepth=1), dict(excl=0), dict(min_len=64), dict(init=0, thr=150, pen=75, bonus=25),
            dict(lazy=0, depth=2, gdepth=2)]


@pytest.mark.parametrize("kw", VARIANTS, ids=lambda k: "-".join(f"{a}{b}" for a, b in k.items()) or "mode1")
def test_decode_parameter_variants(kw):
    files = synth_files(21, 4, 3, 40_000, 8e-3, 8e-4)
    got = _check(encode_collection(files, _pp(**kw)))
    assert got["launches"] > 0


@pytest.mark.parametrize("tier", ["near", "div", "far"])
@pytest.mark.parametrize("g0_is_target", [True, False])
def test_decode_tiers_and_g0_modes(tier, g0_is_target):
    p = dict(near=(1e-4, 1e-5), div=(1e-2, 1e-4), far=(6e-2, 2e-3))[tier]
    files = synth_files(22, 6, 2, 150_000, *p)
    enc = encode_collection(files, _pp(), g0_is_target=g0_is_target, rc_factor=8 if tier != "far" else 128)
    got = _check(enc)
    if tier == "near":
        assert got["batches"] <= 3, "prologue, file 0 (its second contig is new to the reference), then every other target in ONE launch"
    if tier == "far":
        assert got["batches"] > 2


def test_decode_whole_target_reverse_complement():
    files = synth_files(23, 4, 2, 50_000, 1e-4, 1e-5)
    _check(encode_collection(files, _pp()), cir=False)


def test_decode_lock_positions():
    files = synth_files(24, 4, 2, 50_000, 8e-3, 8e-4)
    enc = encode_collection(files, _pp())
    _check(enc, locks=[2**64 - 1] * len(enc.targets))


def test_decode_long_lengths_and_many_small_contigs():
    """match lengths that need the 0xFFFF escape of the frugal stream (> 65534, incl. one whose low half is 0xFFFF), next to a
    file of many short contigs"""
    base = synth_files(25, 1, 1, 400_000, 0, 0)[0][0]
    a = base.copy()
    a[131071 + 5] = ord("A") if a[131071 + 5] != ord("A") else ord("C")   # a match of exactly ... around 2^17
    small = [np.ascontiguousarray(base[i * 700:(i + 1) * 700 + 33]) for i in range(150)]
    files = [[base], [a], small, [base[1000:390_000]]]
    enc = encode_collection(files, _pp())
    got = _check(enc)
    assert max(int(x) for x in got["lens"]) >= 389_000


def test_decode_empty_and_tiny_inputs():
    g = _gpu()
    pp = _pp()
    files = [[np.frombuffer(b"ACGTACGTAC", dtype=np.uint8)], [np.frombuffer(b"ACGTTCGTAC", dtype=np.uint8)]]
    _check(encode_collection(files, pp, limit=1 << 22))
    # no targets at all: only the G0 prologue
    enc = encode_collection([[np.frombuffer(b"ACGT" * 50, dtype=np.uint8)]], pp, g0_is_target=False, limit=1 << 22)
    _check(enc)


def test_decode_rejects_inconsistent_streams():
    import mbgc_b200
    files = synth_files(26, 3, 2, 30_000, 8e-3, 8e-4)
    enc = encode_collection(files, _pp())
    bad = dict(enc.streams)
    bad["gapFlags"] = bad["gapFlags"][: len(bad["gapFlags"]) // 3]
    enc.streams = bad
    with pytest.raises(mbgc_b200.MatcherError):
        _device_decode(enc)
    bad2 = dict(bad)
    enc.streams = bad2
    enc.seq_counts = [c + 5 for c in enc.seq_counts]
    with pytest.raises(mbgc_b200.MatcherError):
        _device_decode(enc)


@pytest.mark.skipif(not have_fixtures(), reason="reference fixtures not available")
@pytest.mark.parametrize("mode", [0, 1, 2])
def test_decode_listeria(mode):
    names = [l.strip() for l in open(os.path.join(FIXTURES, "seqlist.txt")) if l.strip()]
    files = [[np.frombuffer(s, dtype=np.uint8) for _, s in read_fasta(os.path.join(FIXTURES, n))] for n in names]
    pp = OraclePackParams.for_mode(mode)
    enc = encode_collection(files, pp, rc_factor=128 if mode >= 2 else 8, skip_margin=24 if mode >= 2 else 16)
    _check(enc)
