"""Row a12 on the GPU: MBGC_Encoder::processLiteral (mbgccoder/MBGC_Encoder.cpp:128-135) behind
mbgcg_set_literal_ref_ext / mbgcg_pack_literal_ref_ext, against the oracle's restatement (mbo_process_matches_lit, pinned
against stream digests of the reference's developer build run with -N / -F in tests/test_oracle_golden.py)."""
import ctypes as C
import os
import tempfile

import numpy as np
import pytest

from seq_schedule import LitRange
from test_gpu_pack import Pair, _check, _params
from util import LITERAL_EXT_CASES, NO_LOCK, literal_ext_collection, literal_ext_genomes, oracle_lib, u8p

pytestmark = pytest.mark.gpu
STREAMS = ("literals", "mapOff", "mapOff5", "mapLen", "gapDelta", "gapFlags")


def _oracle_pack_lit(o, pp, text, matches, min_len, ctx):
    """mbo_process_matches_lit on fresh streams -> (streams, kept, unmatched, [(start, len)])"""
    lib = oracle_lib()
    lib.mbo_process_matches_lit.restype = C.c_uint64
    st = lib.mbo_streams_create()
    a = np.ascontiguousarray(text, dtype=np.uint8)
    mt = np.ascontiguousarray(matches, dtype=np.uint64).reshape(-1, 3).copy()
    n = C.c_uint64(len(mt))
    buf = mt if len(mt) else np.zeros((1, 3), dtype=np.uint64)
    cap = len(mt) + 2
    rbuf = (LitRange * cap)()
    nr = C.c_uint64(0)
    arr = (C.c_uint64 * 1)()
    unmatched = lib.mbo_process_matches_lit(st, C.byref(pp), o.h, buf.ctypes.data_as(C.c_void_p), C.byref(n),
                                            a.ctypes.data_as(u8p), C.c_uint64(a.size), C.c_uint64(NO_LOCK), arr, C.c_uint64(0),
                                            C.c_uint64(min_len), C.c_uint64(ctx), rbuf, C.c_uint64(cap), C.byref(nr))
    out = {}
    for i, name in enumerate(STREAMS):
        k = lib.mbo_stream_size(st, i)
        p = lib.mbo_stream_data(st, i)
        out[name] = bytes(np.ctypeslib.as_array(C.cast(p, u8p), shape=(k,))) if k else b""
    lib.mbo_streams_destroy(st)
    return out, mt[:n.value].copy(), unmatched, [(int(rbuf[k].start), int(rbuf[k].len)) for k in range(nr.value)]


@pytest.mark.parametrize("case", LITERAL_EXT_CASES + ((1, 0, 0), (1, 5, 1000), (1, 10 ** 9, 8)))
def test_literal_ranges_equal_oracle(case):
    """a batch of genomes against G0 + RC: the ranges of every text, the streams beside them, unmatched counts"""
    mode, min_len, ctx = case
    gs = literal_ext_genomes(7, 120_000, 5)
    pr = Pair(8 << 20, 24 if mode >= 2 else 16)
    pr.load(gs[0])
    po, pd = _params(frugal=0 if mode == 0 else 1)
    targets = gs[1:] + [gs[0][:5000], np.frombuffer(b"ACGTTGCA" * 40, dtype=np.uint8)]  # + an all-matched and an unmatched text
    lists = [pr.o.match(t, 32) for t in targets]
    want, want_ranges = [], []
    for i, (t, mt) in enumerate(zip(targets, lists)):
        s, k, u, r = _oracle_pack_lit(pr.o, po, t, mt, min_len, ctx)
        want.append((s, k, u))
        want_ranges += [(i, a, b) for a, b in r]
    pr.g.setLiteralRefExt(min_len, ctx, 128)
    got = pr.g.packBatch(pd, targets, lists)
    _check(got, want, f"literal ext {case}")
    ranges = pr.g.packLiteralRefExt()
    assert [tuple(int(x) for x in r) for r in ranges] == want_ranges
    if min_len < 10 ** 9:
        assert len(want_ranges) > 0
    # resident route: lists and texts stay on the device between the match call and the pack call
    pr.g.matchBatch(targets, 32, fetch=False)
    got2 = pr.g.packBatch(pd, None, None, lens=[t.size for t in targets])
    _check(got2, want, f"literal ext resident {case}")
    assert [tuple(int(x) for x in r) for r in pr.g.packLiteralRefExt()] == want_ranges
    # off again: no ranges, same streams
    pr.g.setLiteralRefExt(None)
    got3 = pr.g.packBatch(pd, targets, lists)
    _check(got3, want, "literal ext off")
    assert len(pr.g.packLiteralRefExt()) == 0
    pr.close()


def test_literal_ranges_multi_pass(monkeypatch):
    """texts split over several device passes: text indices stay those of the call"""
    monkeypatch.setenv("MBGC_B200_MAX_PASS_BYTES", "250000")
    gs = literal_ext_genomes(6, 100_000, 11)
    pr = Pair(8 << 20)
    pr.load(gs[0])
    po, pd = _params()
    targets = gs[1:]
    lists = [pr.o.match(t, 32) for t in targets]
    want_ranges = []
    for i, (t, mt) in enumerate(zip(targets, lists)):
        _, _, _, r = _oracle_pack_lit(pr.o, po, t, mt, 64, 8)
        want_ranges += [(i, a, b) for a, b in r]
    pr.g.setLiteralRefExt(64, 8, 128)
    pr.g.packBatch(pd, targets, lists)
    assert [tuple(int(x) for x in r) for r in pr.g.packLiteralRefExt()] == want_ranges and len(want_ranges) > 3
    pr.close()


@pytest.mark.parametrize("window", [0, 3])
@pytest.mark.parametrize("case", LITERAL_EXT_CASES)
def test_collection_pass_with_literal_ref_extension(case, window, tmp_path):
    """the sequential schedule of the collection pass (mbgcg_seq_*) with processLiteral on: contigs that are not loaded
    whole extend the reference with their long literals, device to device; streams, reference length and counters equal
    the oracle's schedule, whose streams are those of `mbgc-dev c -t 1 -N .. -F ..` (tests/test_oracle_golden.py)"""
    from test_gpu_collection import check_against_oracle
    mode, min_len, ctx = case
    names = literal_ext_collection(str(tmp_path))
    paths = [os.path.join(str(tmp_path), n) for n in names]
    info = check_against_oracle(paths, mode, [0], 1 << 20, lit_ext=(min_len, ctx), max_window=window)
    assert info["stats"]["extensions"] > 0


def test_collection_pass_literal_ext_two_replicas(tmp_path):
    from test_gpu_collection import check_against_oracle, n_devices
    if n_devices() < 2:
        pytest.skip("needs two GPUs")
    names = literal_ext_collection(str(tmp_path))
    paths = [os.path.join(str(tmp_path), n) for n in names]
    check_against_oracle(paths, 1, [0, 1], 1 << 19, lit_ext=(64, 8))


def test_parallel_schedule_refuses_literal_ext(tmp_path):
    import mbgc_b200
    from mbgc_b200.matcher import SeqConfig
    names = literal_ext_collection(str(tmp_path), n=3, length=50_000)
    g = mbgc_b200.Matcher(8 << 20, 32, 16, 1, 16)
    g.disableSlidingWindow()
    g.setLiteralRefExt(64, 8, 128)
    with pytest.raises(mbgc_b200.MatcherError):
        g.compressSequential([open(os.path.join(str(tmp_path), n), "rb").read() for n in names], SeqConfig.for_mode(1, parallel=True))
    g.close()
