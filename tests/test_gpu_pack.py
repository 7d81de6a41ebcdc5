"""GPU parity of K6 (stream packing): mbgcg_pack_batch against the oracle's restatement of
MBGC_Encoder::processMatches (oracle/mbgc_oracle.c: mbo_process_matches, pinned against the reference's own
stream dumps in tests/test_oracle_golden.py) -- six streams, surviving match list, unmatched count, bit-exact."""
import ctypes as C
import os

import numpy as np
import pytest

from util import (FIXTURES, NO_LOCK, OracleMatcher, PackParams as OraclePackParams, have_fixtures, oracle_lib, read_fasta,
                  synth_base, synth_mutant, u8p)

pytestmark = pytest.mark.gpu

STREAMS = ("literals", "mapOff", "mapOff5", "mapLen", "gapDelta", "gapFlags")
_ACGT = np.frombuffer(b"ACGT", dtype=np.uint8)


def _params(enable=1, min_len=256, depth=64, gdepth=2, lazy=1, frugal=1, bit40=0, excl=1, init=125, thr=500, pen=50, bonus=50):
    """-> (oracle params, device params) with the same values"""
    from mbgc_b200.matcher import PackParams
    o = OraclePackParams(enable, min_len, depth, gdepth, lazy, frugal, bit40, excl, init, thr, pen, bonus, 32, 128)
    d = PackParams(enable, depth, gdepth, lazy, frugal, bit40, excl, init, thr, pen, bonus, 0, min_len)
    return o, d


def _oracle_pack(o, pp, text, matches, lock=NO_LOCK, loaded=()):
    """one mbo_process_matches call on fresh streams -> (dict of bytes, kept matches, unmatched)"""
    lib = oracle_lib()
    st = lib.mbo_streams_create()
    a = np.ascontiguousarray(text, dtype=np.uint8)
    mt = np.ascontiguousarray(matches, dtype=np.uint64).reshape(-1, 3).copy()
    n = C.c_uint64(len(mt))
    arr = (C.c_uint64 * max(len(loaded), 1))(*loaded)
    buf = mt if len(mt) else np.zeros((1, 3), dtype=np.uint64)
    tptr = a.ctypes.data_as(u8p) if a.size else C.cast(C.create_string_buffer(8), u8p)
    unmatched = lib.mbo_process_matches(st, C.byref(pp), o.h, buf.ctypes.data_as(C.c_void_p), C.byref(n), tptr, a.size, lock,
                                        arr, len(loaded))
    out = {}
    for i, name in enumerate(STREAMS):
        k = lib.mbo_stream_size(st, i)
        p = lib.mbo_stream_data(st, i)
        out[name] = bytes(np.ctypeslib.as_array(C.cast(p, u8p), shape=(k,))) if k else b""
    lib.mbo_streams_destroy(st)
    return out, mt[:n.value].copy(), unmatched


def _first_diff(a, b):
    n = min(len(a), len(b))
    for i in range(n):
        if a[i] != b[i]:
            return f"first difference at byte {i} of {len(a)} / {len(b)}: cuda {a[max(0, i - 4):i + 8].hex()} oracle {b[max(0, i - 4):i + 8].hex()}"
    return f"lengths differ: cuda {len(a)} oracle {len(b)}"


def _check(got, want_list, tag=""):
    """got: packBatch result; want_list: [(streams, kept, unmatched)] per text"""
    for i, (ws, wk, wu) in enumerate(want_list):
        for name in STREAMS:
            assert got[name][i] == ws[name], f"{tag} text {i} stream {name}: " + _first_diff(got[name][i], ws[name])
        assert int(got["unmatched"][i]) == int(wu), f"{tag} text {i}: unmatched {got['unmatched'][i]} vs {wu}"
        assert int(got["kept"][i]) == len(wk), f"{tag} text {i}: kept {got['kept'][i]} vs {len(wk)}"


class Pair:
    """oracle matcher + CUDA matcher holding the same reference"""

    def __init__(self, limit, skip_margin=16):
        import mbgc_b200
        self.o = OracleMatcher(limit, 32, 16, 1, skip_margin)
        self.g = mbgc_b200.Matcher(limit, 32, 16, 1, skip_margin)
        self.o.disable_sliding_window()
        self.g.disableSlidingWindow()
        self.loaded = []

    def load(self, text, rc=True):
        self.o.load_ref(text, rc=rc, add_sep=True)
        self.g.loadRef(text, rc, True, 0)

    def sep(self):
        self.o.load_separator(0)
        self.g.loadSeparator(0)

    def close(self):
        self.g.close()
        self.o.close()


def _run_all_routes(pr, texts, po, pd, locks=None, loaded=(), tag=""):
    """oracle per text vs: one host-path call per text, one host-path batch, the resident route"""
    o, g = pr.o, pr.g
    lk = locks if locks is not None else [NO_LOCK] * len(texts)
    lists = [o.match(t, 32, l) for t, l in zip(texts, lk)]
    want = [_oracle_pack(o, po, t, m, l, loaded) for t, m, l in zip(texts, lists, lk)]
    # (a) batch, host match lists
    got = g.packBatch(pd, texts, lists, lockPos=locks, loadedPosArr=loaded)
    _check(got, want, tag + " host-batch")
    kept = g.packKeptMatches()
    assert np.array_equal(kept, np.concatenate([w[1] for w in want]) if want else kept), tag + " kept lists differ"
    # (b) resident: lists and texts of the preceding match call
    glists = g.matchBatch(texts, 32, locks)
    for a, b in zip(glists, lists):
        assert np.array_equal(a, b), tag + " match lists differ (matcher parity)"
    got = g.packBatch(pd, texts, None, lockPos=locks, loadedPosArr=loaded)
    _check(got, want, tag + " resident")
    # (b') the same with the lists never leaving the device (out == NULL in the match call)
    cnts = g.matchBatch(texts, 32, locks, fetch=False)
    assert [int(c) for c in cnts] == [len(m) for m in lists]
    got = g.packBatch(pd, texts, None, lockPos=locks, loadedPosArr=loaded)
    _check(got, want, tag + " resident, lists on the device only")
    # (c) one call per text
    for i, t in enumerate(texts[:6]):
        got = g.packBatch(pd, [t], [lists[i]], lockPos=[lk[i]] if locks is not None else None, loadedPosArr=loaded)
        _check(got, [want[i]], tag + f" single {i}")
    return want, g.packTiming()


def _collection(seed, n, length, p_snp, p_indel):
    base = synth_base(length, seed)
    return base, [synth_mutant(base, seed * 7919 + i, p_snp, p_indel) for i in range(1, n)]


@pytest.mark.parametrize("mode_frugal", [0, 1])
@pytest.mark.parametrize("tier", ["near", "div"])
def test_pack_matches_oracle(tier, mode_frugal):
    p_snp, p_indel = (1e-3, 1e-4) if tier == "near" else (1e-2, 1e-3)
    base, texts = _collection(101, 5, 200_000, p_snp, p_indel)
    pr = Pair(1 << 23)
    try:
        pr.load(base)
        loaded = [pr.o.loading_position]
        po, pd = _params(frugal=mode_frugal)
        want, tm = _run_all_routes(pr, texts, po, pd, loaded=loaded, tag=tier)
        assert tm["launches"] > 0
        assert sum(len(w[0]["gapFlags"]) for w in want) > 0
    finally:
        pr.close()


@pytest.mark.parametrize("kw", [dict(lazy=0), dict(enable=0), dict(depth=0), dict(depth=3), dict(depth=127), dict(gdepth=0),
                                dict(gdepth=1), dict(bit40=1), dict(excl=0), dict(min_len=64), dict(min_len=100000),
                                dict(init=0, thr=150, pen=75, bonus=25), dict(lazy=0, depth=2, gdepth=2)],
                         ids=lambda k: "-".join(f"{a}{b}" for a, b in k.items()))
def test_pack_parameter_variants(kw):
    base, texts = _collection(202, 4, 120_000, 8e-3, 8e-4)
    pr = Pair(1 << 22)
    try:
        pr.load(base)
        po, pd = _params(**kw)
        _run_all_routes(pr, texts, po, pd, loaded=[pr.o.loading_position], tag=str(kw))
    finally:
        pr.close()


def test_pack_regions_and_locks():
    """several loaded regions (separators, refExtLoadedPosArr with more than one entry: the lazy pairing rule), lock
    positions inside the reference (pairedWith(rhs, pairBreakSrcPos), the left-extension guard)"""
    base, texts = _collection(303, 6, 90_000, 6e-3, 6e-4)
    pr = Pair(1 << 22)
    try:
        pr.load(base)
        loaded = [pr.o.loading_position]
        for t in texts[:3]:
            start = pr.o.loaded_ref_length
            pr.load(t, rc=False)
            pr.sep()
            loaded.append(loaded[-1] + (pr.o.loaded_ref_length - start))
        po, pd = _params()
        # texts stitched from pieces of different regions so that diagonals cross region boundaries
        mixed = [np.concatenate([texts[0][:30_000], texts[1][30_000:60_000], base[60_000:]]),
                 np.concatenate([texts[2][10_000:50_000], texts[0][50_000:80_000]]), texts[4], texts[3]]
        _run_all_routes(pr, mixed, po, pd, loaded=loaded, tag="regions")
        ref_len = pr.o.ref_length
        locks = [ref_len // 3, 90_000, ref_len - 1000, 1]
        _run_all_routes(pr, mixed, po, pd, locks=locks, loaded=loaded, tag="locks")
        po0, pd0 = _params(lazy=0)
        _run_all_routes(pr, mixed, po0, pd0, locks=locks, loaded=loaded, tag="locks-nolazy")
    finally:
        pr.close()


def test_pack_pathological_texts():
    """N runs, lower case, homopolymers, tandem repeats (gap-breaking matches), empty / tiny / unmatched texts, long
    literal runs (the long-run copy kernel) at odd alignments"""
    r = np.random.default_rng(7)
    unit = _ACGT[r.integers(0, 4, 37)]
    tandem = np.tile(unit, 400)
    rnd = _ACGT[r.integers(0, 4, 60_000)]
    novel = _ACGT[r.integers(0, 4, 50_001)]
    polyA = np.full(5_000, ord("A"), dtype=np.uint8)
    nrun = np.full(3_000, ord("N"), dtype=np.uint8)
    lower = np.frombuffer(bytes(rnd[:4000]).lower(), dtype=np.uint8)
    ref = np.concatenate([rnd, polyA, rnd[:10_000], nrun, tandem, lower, rnd[5_000:25_000]])
    t1 = np.concatenate([rnd[100:20_000], nrun[:1500], polyA[:2000], tandem[:5000], rnd[:3000], lower[:2000]])
    t2 = np.concatenate([tandem[11:9000], polyA, rnd[20_000:], nrun, rnd[:50]])
    dup = np.concatenate([rnd[:8000], rnd[:8000], rnd[4000:12000]])
    t3 = np.concatenate([novel[:4097], rnd[1000:3000], novel[4097:30_001], rnd[7:2007], novel[30_001:]])
    snp = rnd[:40_000].copy()
    snp[np.arange(50, 40_000, 61)] = ord("N")          # dense mismatches: extensions chain through whole gaps
    texts = [t1, t2, dup, t3, novel, snp, rnd[:31], rnd[:32], rnd[:33], np.zeros(0, np.uint8), rnd[:5], polyA[:100]]
    pr = Pair(1 << 22)
    try:
        pr.load(ref)
        for kw in (dict(), dict(frugal=0, lazy=0), dict(min_len=1000)):
            po, pd = _params(**kw)
            want, _ = _run_all_routes(pr, texts, po, pd, loaded=[pr.o.loading_position], tag=f"patho{kw}")
        # texts matched right at the loading position: the right-extension guards (:322-333)
        tail = ref[-5000:].copy()
        tail2 = np.concatenate([ref[-3000:], novel[:200]])
        po, pd = _params()
        _run_all_routes(pr, [tail, tail2], po, pd, loaded=[pr.o.loading_position], tag="tail")
    finally:
        pr.close()


def test_pack_many_small_contigs_and_passes(monkeypatch):
    r = np.random.default_rng(21)
    base = synth_base(150_000, 21)
    texts = []
    for i in range(300):
        ln = int(r.integers(20, 3000))
        st = int(r.integers(0, base.size - ln))
        t = base[st:st + ln].copy()
        k = int(r.integers(0, 4))
        if k and ln > 40:
            t[r.integers(0, ln, k)] = ord("A")
        texts.append(t)
    pr = Pair(1 << 22)
    try:
        pr.load(base)
        po, pd = _params()
        o, g = pr.o, pr.g
        lists = [o.match(t, 32) for t in texts]
        want = [_oracle_pack(o, po, t, m, NO_LOCK, [o.loading_position]) for t, m in zip(texts, lists)]
        got = g.packBatch(pd, texts, lists, loadedPosArr=[o.loading_position])
        _check(got, want, "small")
        monkeypatch.setenv("MBGC_B200_MAX_PASS_BYTES", "40000")   # many passes: streams concatenated on the host
        got = g.packBatch(pd, texts, lists, loadedPosArr=[o.loading_position])
        _check(got, want, "small multi-pass")
        kept = g.packKeptMatches()
        assert np.array_equal(kept, np.concatenate([w[1] for w in want]))
        monkeypatch.delenv("MBGC_B200_MAX_PASS_BYTES")
    finally:
        pr.close()


@pytest.mark.skipif(not have_fixtures(), reason="reference fixtures not built")
def test_pack_listeria_sequential_schedule():
    """the reference's deterministic (-t 1) schedule on the shipped Listeria example, matcher and packer both on the
    device (resident route), against the oracle-only emulation (which reproduces the reference's stream dumps,
    tests/test_oracle_golden.py): every stream identical, 1 418 gap-breaking matches removed on the way"""
    from seq_schedule import SEQ_SEPARATOR_MARK, encode_streams, reference_factor
    import mbgc_b200
    paths = [os.path.join(FIXTURES, l.strip()) for l in open(os.path.join(FIXTURES, "seqlist.txt")) if l.strip()]
    for mode in (1, 0):
        want, stats = encode_streams(paths, mode)
        recs = [read_fasta(p) for p in paths]
        c0 = recs[0][0][1]
        basic = max(os.path.getsize(paths[0]), 1 << 21)
        limit = max(min(reference_factor(len(paths)) * basic * 2, 2**32 - 1), len(c0))
        _, pd = _params(frugal=0 if mode == 0 else 1)
        g = mbgc_b200.Matcher(limit, 32, 16, 1, 16)
        try:
            g.disableSlidingWindow()
            g.loadRef(c0, True, True, 0)
            loaded = [g.getLoadingPosition()]
            out = {n: [] for n in STREAMS}
            out["literals"].append(bytes(c0) + bytes([SEQ_SEPARATOR_MARK]))
            removed = 0
            for recs_i in recs:
                start = g.getLoadedRefLength()
                for _, seq in recs_i:
                    a = np.frombuffer(seq, dtype=np.uint8)
                    mt = g.matchTexts(a, 32)
                    res = g.packBatch(pd, [a], None, loadedPosArr=loaded)
                    removed += len(mt) - int(res["kept"][0])
                    for n in STREAMS:
                        out[n].append(res[n][0])
                    unmatched = int(res["unmatched"][0])
                    if unmatched * 128 > a.size:
                        g.loadRef(a, unmatched * 8 > a.size, True, 0)
                    out["literals"].append(bytes([SEQ_SEPARATOR_MARK]))
                out["gapFlags"].append(bytes([(ord(';') + 128) & 0xFF]))
                g.loadSeparator(0)
                loaded.append(loaded[-1] + (g.getLoadedRefLength() - start))
            for n in STREAMS:
                got = b"".join(out[n])
                assert got == want[n], f"mode {mode} stream {n}: " + _first_diff(got, want[n])
            if mode == 1:
                assert removed == 1418  # SURVEY.md 8c dev counter of the reference
        finally:
            g.close()


def test_pack_error_paths(monkeypatch):
    """what the device packer refuses, and how: parameters outside its scope (MBGCG_E_UNSUPPORTED), a resident call
    without the match call it refers to (MBGCG_E_PARAM), lists asked to stay on the device for a multi-pass batch"""
    import mbgc_b200
    from mbgc_b200.matcher import MatcherError
    base, texts = _collection(404, 3, 60_000, 1e-3, 1e-4)
    pr = Pair(1 << 22)
    try:
        pr.load(base)
        g = pr.g
        _, pd = _params(gdepth=3)
        with pytest.raises(MatcherError) as e:
            g.packBatch(pd, texts, [g.matchTexts(t) for t in texts])
        assert e.value.code == -3
        _, pd = _params(depth=128)
        with pytest.raises(MatcherError) as e:
            g.packBatch(pd, texts, [g.matchTexts(t) for t in texts])
        assert e.value.code == -3
        _, pd = _params()
        g.matchBatch(texts, 32)
        with pytest.raises(MatcherError) as e:          # other texts than those of the match call
            g.packBatch(pd, texts[:1], None)
        assert e.value.code == -2
        g.packBatch(pd, texts, [g.matchTexts(t) for t in texts])   # a host-route call overwrites the device texts ...
        with pytest.raises(MatcherError) as e:                      # ... so nothing is resident any more
            g.packBatch(pd, texts, None)
        assert e.value.code == -2
        monkeypatch.setenv("MBGC_B200_MAX_PASS_BYTES", "70000")
        with pytest.raises(MatcherError) as e:
            g.matchBatch(texts, 32, fetch=False)
        assert e.value.code == -2
        lists = g.matchBatch(texts, 32)                              # with the lists fetched, two passes are fine
        monkeypatch.delenv("MBGC_B200_MAX_PASS_BYTES")
        for a, t in zip(lists, texts):
            assert np.array_equal(a, pr.o.match(t))
    finally:
        pr.close()


def test_pack_fixed_point_rerun_path(monkeypatch):
    """starting with a single sweep leaves the convergence flag set whenever anything pairs or is removed: the host must
    notice, run again with more sweeps, and arrive at the same streams"""
    base, texts = _collection(505, 4, 150_000, 5e-3, 5e-4)
    pr = Pair(1 << 22)
    try:
        pr.load(base)
        po, pd = _params()
        lists = [pr.o.match(t) for t in texts]
        want = [_oracle_pack(pr.o, po, t, m, NO_LOCK, [pr.o.loading_position]) for t, m in zip(texts, lists)]
        monkeypatch.setenv("MBGC_B200_PACK_SWEEPS", "1")
        got = pr.g.packBatch(pd, texts, lists, loadedPosArr=[pr.o.loading_position])
        _check(got, want, "one sweep first")
        assert pr.g.packTiming()["sweeps"][1] >= 2
        monkeypatch.delenv("MBGC_B200_PACK_SWEEPS")
    finally:
        pr.close()
