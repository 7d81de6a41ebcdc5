"""The CPU oracle against the golden vectors generated FROM THE REFERENCE (tests/golden/make_golden.py):
match lists, table contents and cursor state of every scenario; the six encoder streams of the deterministic
schedule; hash known answers.  Needs neither /root/reference nor oracle/_ref at run time."""
import hashlib
import json
import os
import tempfile

import numpy as np
import pytest

from scenarios import all_scenarios, replay
from seq_schedule import encode_streams
from util import FIXTURES, GOLDEN, LITERAL_EXT_CASES, OracleMatcher, have_fixtures, literal_ext_collection, oracle_lib, ptr

G = json.load(open(os.path.join(GOLDEN, "golden.json")))


def digest(arrs):
    h = hashlib.sha256()
    for a in arrs:
        h.update(np.ascontiguousarray(a, dtype=np.uint64).tobytes())
        h.update(b"|")
    return h.hexdigest()


@pytest.mark.parametrize("skip_margin", [16, 24])
@pytest.mark.parametrize("sc", all_scenarios(), ids=lambda s: s["name"])
def test_scenarios(sc, skip_margin):
    want = G["scenarios"][f"{sc['name']}/{skip_margin}"]
    o = OracleMatcher(sc["limit"], 32, 16, 1, skip_margin)
    res, state = replay(o, sc)
    assert [len(x) for x in res] == want["counts"]
    assert digest(res) == want["matches"]
    assert [int(x) for x in state] == want["state"]
    assert hashlib.sha256(o.table().tobytes()).hexdigest() == want["table"]
    o.close()


def test_prefilter_is_redundant():
    """the l/r 4-byte side-probe filter (ref .cpp:203-204,223-226) never changes a result: the CUDA path omits it"""
    for sc in all_scenarios()[:6]:
        a = OracleMatcher(sc["limit"])
        b = OracleMatcher(sc["limit"])
        ra, _ = replay(a, sc)
        orig = b.match
        b.match = lambda t, ml=32, lock=2**64 - 1: orig(t, ml, lock, prefilter=False)
        rb, _ = replay(b, sc)
        for x, y in zip(ra, rb):
            assert np.array_equal(x, y)


def test_hash_known_answers():
    """every slot the reference wrote for a seeded text: slot = hash28(k-mer) & mask, value = pos >> 4"""
    lib = oracle_lib()
    h = G["hash"]
    rng = np.random.default_rng(h["text_seed"])
    text = np.frombuffer(b"ACGT", dtype=np.uint8)[rng.integers(0, 4, h["text_len"])]
    ref = np.concatenate([np.zeros(1, np.uint8), text])  # reference position 1 = text[0]
    mask = lib.mbo_hash_size(1 << 20, 16) - 1
    want = {s: v for s, v in h["slots"]}
    got = {}
    for pos in range(16, text.size + 1 - 28 + 1, 16):
        got[lib.mbo_hash(ptr(np.ascontiguousarray(ref[pos:pos + 28])), 28) & mask] = pos >> 4
    assert got == want


def _golden_collection(d):
    from mbgc_b200.synth import synth_collection, write_fasta
    gs = synth_collection(6, 200_000, 99, 5e-3, 5e-4)
    paths = []
    for i, g in enumerate(gs):
        p = os.path.join(d, f"g{i:04d}.fna")
        write_fasta(p, f"syn_{i:04d} synthetic E. coli-like genome {i}", g)
        paths.append(p)
    return paths


@pytest.mark.parametrize("mode", [0, 1, 2])
def test_streams_of_sequential_schedule(mode):
    """matchTexts + processMatches + extension decisions, replayed with the oracle, give the reference's streams"""
    want = G["streams"][str(mode)]
    with tempfile.TemporaryDirectory() as d:
        out, _ = encode_streams(_golden_collection(d), mode)
    for name in ("literals", "gapDelta", "gapFlags", "mapOff", "mapLen"):
        assert len(out[name]) == want[name]["size"], name
        assert hashlib.sha256(out[name]).hexdigest() == want[name]["sha256"], name


@pytest.mark.parametrize("case", LITERAL_EXT_CASES)
def test_streams_with_literal_ref_extension(case):
    """row a12: processLiteral (MBGC_Encoder.cpp:128-135) switched on as the reference's developer build does with -N / -F;
    the oracle's schedule, extending the reference with the long literals of contigs that are not loaded whole, gives
    the streams of `mbgc-dev c -t 1 -N .. -F ..`"""
    mode, min_len, ctx = case
    want = G["literal_ext"][f"{mode}:{min_len}:{ctx}"]
    with tempfile.TemporaryDirectory() as d:
        names = literal_ext_collection(d)
        out, stats = encode_streams([os.path.join(d, n) for n in names], mode, lit_ext=(min_len, ctx))
    assert sum(len(r) for r in stats["lit_ranges"]) > 0
    for name in ("literals", "gapDelta", "gapFlags", "mapOff", "mapLen"):
        assert len(out[name]) == want[name]["size"], name
        assert hashlib.sha256(out[name]).hexdigest() == want[name]["sha256"], name


@pytest.mark.skipif(not have_fixtures(), reason="reference fixtures not present (oracle/_ref/fixtures)")
@pytest.mark.parametrize("mode", [0, 1, 2])
def test_streams_listeria(mode):
    want = G["listeria"][str(mode)]
    paths = [os.path.join(FIXTURES, l.strip()) for l in open(os.path.join(FIXTURES, "seqlist.txt"))]
    out, stats = encode_streams(paths, mode)
    for name in ("literals", "gapDelta", "gapFlags", "mapOff", "mapLen"):
        assert hashlib.sha256(out[name]).hexdigest() == want[name]["sha256"], name
    if mode == 1:  # the reference's own dev counters
        assert stats["matches"] == 29731 and stats["unmatched"] == 5009573 and stats["ref_length"] == 16584539


def test_literal_ranges_properties():
    """row a12, oracle alone: with the option off the call is the plain processMatches; with it on, every range lies inside
    its text, covers a literal run longer than the threshold plus at most `context` bytes on either side, and a contig that
    IS proper for a whole-contig extension gives none (MBGC_Encoder.cpp:211)"""
    import ctypes as C
    from seq_schedule import LitRange
    from util import NO_LOCK, PackParams, literal_ext_genomes, u8p
    lib = oracle_lib()
    lib.mbo_process_matches_lit.restype = C.c_uint64
    gs = literal_ext_genomes(4, 60_000, 3)
    o = OracleMatcher(1 << 22, 32, 16, 1, 16)
    o.disable_sliding_window()
    o.load_ref(gs[0], rc=True, add_sep=True)

    def run(t, min_len, ctx, factor):
        pp = PackParams.for_mode(1)
        pp.unmatchedFractionFactor = factor
        mt = np.ascontiguousarray(o.match(t, 32))
        st = lib.mbo_streams_create()
        n = C.c_uint64(len(mt))
        rb = (LitRange * (len(mt) + 2))()
        nr = C.c_uint64(0)
        arr = (C.c_uint64 * 1)()
        a = np.ascontiguousarray(t, dtype=np.uint8)
        buf = mt if len(mt) else np.zeros((1, 3), dtype=np.uint64)
        u = lib.mbo_process_matches_lit(st, C.byref(pp), o.h, buf.ctypes.data_as(C.c_void_p), C.byref(n), a.ctypes.data_as(u8p),
                                        C.c_uint64(a.size), C.c_uint64(NO_LOCK), arr, C.c_uint64(0), C.c_uint64(min_len),
                                        C.c_uint64(ctx), rb, C.c_uint64(len(mt) + 2), C.byref(nr))
        streams = []
        for i in range(6):
            k = lib.mbo_stream_size(st, i)
            p = lib.mbo_stream_data(st, i)
            streams.append(bytes(np.ctypeslib.as_array(C.cast(p, u8p), shape=(k,))) if k else b"")
        lib.mbo_streams_destroy(st)
        return u, streams, [(int(rb[k].start), int(rb[k].len)) for k in range(nr.value)]

    for t in gs[1:]:
        u0, s0, r0 = run(t, 2**64 - 1, 0, 16)       # off
        u1, s1, r1 = run(t, 64, 8, 16)              # factor 16: ~700 unmatched bases of 60 k do not make the contig proper
        u2, s2, r2 = run(t, 64, 0, 16)
        assert r0 == [] and (u0, s0) == (u1, s1) == (u2, s2), "the option must not change the streams"
        assert len(r1) == len(r2) > 0
        for (a1, l1), (a2, l2) in zip(r1, r2):
            assert l2 > 64 and a2 + l2 <= t.size                       # the bare literal run
            assert a1 <= a2 and a2 - a1 <= 8 and a1 + l1 >= a2 + l2 and (a1 + l1) - (a2 + l2) <= 8 and a1 + l1 <= t.size
        # proper for a whole-contig extension (unmatched * factor > length): processLiteral is not called at all
        assert run(t, 64, 8, 1 << 20)[2] == []
    o.close()
