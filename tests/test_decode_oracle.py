"""Row f4: the oracle's restatement of the decoder's match-copy path (oracle/mbgc_oracle_dec.c) against the REFERENCE's own
MBGC_Decoder<false> members driven on the same streams (oracle/_ref/libref_decoder.so: decodeTarget, initReference,
decodeReference, loadRef), and by round trip -- the streams of the (pinned) encoder restatement decode to the bytes that
were encoded.  CPU only."""
import ctypes as C
import os

import numpy as np
import pytest

from decode_cases import (dec_oracle, dec_params_array, encode_collection, oracle_decode, ref_decoder, reference_decode,
                          synth_files)
from util import FIXTURES, PackParams, have_fixtures, read_fasta, u8p

REF = ref_decoder()
needs_ref = pytest.mark.skipif(REF is None, reason="oracle/_ref/libref_decoder.so not built (no /root/reference here)")


def _pp(enable=1, min_len=256, depth=64, gdepth=2, lazy=1, frugal=1, bit40=0, excl=1, init=125, thr=500, pen=50, bonus=50):
    return PackParams(enable, min_len, depth, gdepth, lazy, frugal, bit40, excl, init, thr, pen, bonus, 32, 128)


VARIANTS = [dict(), dict(frugal=0), dict(lazy=0), dict(enable=0), dict(depth=0), dict(depth=3), dict(depth=127), dict(gdepth=0),
            dict(gdepth=1), dict(excl=0), dict(min_len=64), dict(init=0, thr=150, pen=75, bonus=25),
            dict(lazy=0, depth=2, gdepth=2)]


def _check_round_trip(enc, res, g0=None):
    for (f, cnt), (parts, unm) in zip(enc.targets, res):
        assert len(parts) == cnt
        for c, (got, want) in enumerate(zip(parts, enc.contigs[f])):
            assert np.array_equal(got, want), f"file {f} contig {c}: decoded bytes differ from the encoded contig"
    if g0 is not None:
        for got, want in zip(g0, enc.contigs[0]):
            assert np.array_equal(got, want)


def _check_against_reference(enc, locks=None, cir=True):
    res, g0, ref, ref_pos, cur = oracle_decode(enc, locks, cir)
    rres, rref, rref_pos, rcur = reference_decode(REF, enc, locks, cir)
    for t, ((parts, _), rparts) in enumerate(zip(res, rres)):
        assert len(parts) == len(rparts)
        for c, (a, b) in enumerate(zip(parts, rparts)):
            assert np.array_equal(a, b), f"target {t} contig {c}: oracle and reference decoders differ"
    assert ref_pos == rref_pos
    assert cur == rcur, f"stream cursors differ: {cur} vs {rcur}"
    assert np.array_equal(ref, rref), "reference buffers differ"
    return res, g0


@pytest.mark.parametrize("kw", VARIANTS, ids=lambda k: "-".join(f"{a}{b}" for a, b in k.items()) or "mode1")
def test_round_trip_parameter_variants(kw):
    files = synth_files(11, 4, 3, 40_000, 8e-3, 8e-4)
    enc = encode_collection(files, _pp(**kw))
    res, g0, ref, ref_pos, cur = oracle_decode(enc)
    _check_round_trip(enc, res)
    assert cur[0] == len(enc.streams["literals"]) and cur[5] == len(enc.streams["gapFlags"])
    assert cur[1] * 4 == len(enc.streams["mapOff"]) and cur[4] == len(enc.streams["gapDelta"])
    assert ref_pos == enc.ref_total


@needs_ref
@pytest.mark.parametrize("kw", VARIANTS, ids=lambda k: "-".join(f"{a}{b}" for a, b in k.items()) or "mode1")
def test_equals_reference_decoder(kw):
    files = synth_files(12, 4, 2, 50_000, 8e-3, 8e-4)
    enc = encode_collection(files, _pp(**kw))
    res, _ = _check_against_reference(enc)
    _check_round_trip(enc, res)


@needs_ref
@pytest.mark.parametrize("tier", ["near", "div", "far"])
@pytest.mark.parametrize("g0_is_target", [True, False])
def test_tiers_and_g0_modes(tier, g0_is_target):
    p = dict(near=(1e-4, 1e-5), div=(1e-2, 1e-4), far=(6e-2, 2e-3))[tier]
    files = synth_files(13, 5, 2, 60_000, *p)
    enc = encode_collection(files, _pp(), g0_is_target=g0_is_target, rc_factor=8 if tier != "far" else 128)
    res, g0 = _check_against_reference(enc)
    _check_round_trip(enc, res, g0)
    if tier == "far":
        assert enc.ref_total > 4 * 120_000, "the far tier is meant to extend the reference"


@needs_ref
def test_whole_target_reverse_complement_and_no_rc():
    """contigsIndivduallyReversed = false (one reverse complement per target, :627-635) and RC switched off: oracle and
    reference agree byte for byte (the encoder emulation does not know these layouts, so there is no round trip: a
    near-identical collection, which never extends the reference, keeps the streams meaningful)"""
    files = synth_files(14, 4, 2, 50_000, 1e-4, 1e-5)
    enc = encode_collection(files, _pp())
    res, _ = _check_against_reference(enc, cir=False)
    _check_round_trip(enc, res)


@needs_ref
def test_lock_positions_inside_the_reference():
    """explicit lock positions (matchingLocksPosArr): the left-extension guard (:469-472) and loadRef's limit"""
    files = synth_files(15, 4, 2, 50_000, 8e-3, 8e-4)
    enc = encode_collection(files, _pp())
    locks = [2**64 - 1] * len(enc.targets)
    res, _ = _check_against_reference(enc, locks=locks)
    _check_round_trip(enc, res)


@needs_ref
@pytest.mark.parametrize("lazy", [0, 1])
def test_load_ref_wraps_and_locks(lazy):
    """loadRef (:651-675) on its own: truncation at the lock, the wrap to REF_SHIFT, the separator written in front of the lock"""
    L = dec_oracle()
    rng = np.random.default_rng(5)
    total = 10000   # G0 + its reverse complement take 6000 of it
    enc = encode_collection(synth_files(16, 2, 1, 3000, 1e-3, 1e-4), _pp(lazy=lazy))
    enc.ref_total = total
    from decode_cases import _streams_struct, STREAMS, _ptr
    s, keep = _streams_struct(enc)
    d = L.mbo_dec_create(C.byref(enc.pp), total, 1, 1, C.byref(s))
    args = []
    for name in STREAMS:
        a, p = _ptr(enc.streams[name])
        keep.append(a)
        args += [p, len(enc.streams[name])]
    counts = (C.c_uint32 * 2)(1, 1)
    fac = np.array([128, 8], dtype=np.uint8)
    h = REF.refd_create(dec_params_array(enc.pp), total, *args, counts, fac.ctypes.data_as(u8p), None, 1)
    try:
        L.mbo_dec_init_reference(d)
        REF.refd_init_reference(h)
        steps = [(700, total, 0), (900, total, 1), (1500, total, 0), (1800, 2500, 0), (900, 2500, 1), (1200, 2500, 0), (100, 3000, 1),
                 (2000, 4000, 0), (3000, 1500, 1), (7000, 800, 0), (500, 9000, 1), (9000, 9000, 0)]
        for n, lock, rc in steps:
            text = rng.integers(0, 4, n).astype(np.uint8)
            text = np.frombuffer(b"ACGT", dtype=np.uint8)[text].copy()
            L.mbo_dec_load_ref(d, text.ctypes.data_as(u8p), n, lock, rc)
            REF.refd_load_ref(h, text.ctypes.data_as(u8p), n, lock, rc)
            a = np.ctypeslib.as_array(C.cast(L.mbo_dec_ref(d), u8p), shape=(total,))
            b = np.ctypeslib.as_array(C.cast(REF.refd_ref(h), u8p), shape=(total,))
            assert L.mbo_dec_ref_pos(d) == REF.refd_ref_pos(h), (n, lock, rc)
            assert L.mbo_dec_reached_count(d) == REF.refd_reached_count(h)
            assert np.array_equal(a, b), (n, lock, rc)
    finally:
        L.mbo_dec_destroy(d)
        REF.refd_destroy(h)


@pytest.mark.skipif(not have_fixtures(), reason="reference fixtures not available")
@pytest.mark.parametrize("mode", [0, 1, 2])
def test_listeria_round_trip(mode):
    """the shipped example (16-27 contigs per file, N runs): encode with the -t 1 schedule, decode, compare with the FASTA"""
    names = [l.strip() for l in open(os.path.join(FIXTURES, "seqlist.txt")) if l.strip()]
    files = [[np.frombuffer(s, dtype=np.uint8) for _, s in read_fasta(os.path.join(FIXTURES, n))] for n in names]
    pp = PackParams.for_mode(mode)
    enc = encode_collection(files, pp, rc_factor=128 if mode >= 2 else 8, skip_margin=24 if mode >= 2 else 16)
    if REF is not None:
        res, _ = _check_against_reference(enc)
    else:
        res = oracle_decode(enc)[0]
    _check_round_trip(enc, res)


def test_code2mismatch_inverts_mismatch2code():
    from util import oracle_lib
    L = dec_oracle()
    O = oracle_lib()
    for a in b"ACGTN":
        for m in b"ACGTNRYacgt":
            code = O.mbo_mismatch2code(a, m)
            assert L.mbo_code2mismatch(a, code) == m
