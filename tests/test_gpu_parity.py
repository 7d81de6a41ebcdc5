"""GPU parity: the CUDA matcher (through the C ABI, mbgc_b200.Matcher) against the CPU oracle on the same
seeded inputs -- match lists, table contents, reference buffer and cursor state, bit-exact."""
import numpy as np
import pytest

from scenarios import CudaFace, all_scenarios, replay
from util import (FIXTURES, OracleMatcher, have_fixtures, mgmp_ref_limit, read_fasta, synth_base, synth_collection,
                  synth_mutant)

pytestmark = pytest.mark.gpu


def _diff(a, b):
    n = min(len(a), len(b))
    for i in range(n):
        if not np.array_equal(a[i], b[i]):
            return f"first difference at match {i}: cuda {a[i].tolist()} oracle {b[i].tolist()} (counts {len(a)} vs {len(b)})"
    return f"counts differ: cuda {len(a)} oracle {len(b)}"


@pytest.mark.parametrize("skip_margin", [16, 24])
@pytest.mark.parametrize("sc", all_scenarios(), ids=lambda s: s["name"])
def test_scenarios_match_oracle(sc, skip_margin):
    o = OracleMatcher(sc["limit"], 32, 16, 1, skip_margin)
    g = CudaFace(sc["limit"], 32, 16, 1, skip_margin)
    try:
        want, sw = replay(o, sc)
        got, sg = replay(g, sc)
        assert sg == sw
        assert len(got) == len(want)
        for k, (a, b) in enumerate(zip(got, want)):
            assert np.array_equal(a, b), f"op {k}: " + _diff(a, b)
        assert np.array_equal(g.table(), o.table()), "hash table differs"
        n = o.ref_length
        assert np.array_equal(g.ref(n), o.ref(n)), "host mirror of the reference differs"
        assert np.array_equal(g.m.deviceRef(0, n), o.ref(n)), "device reference differs"
    finally:
        g.close()
        o.close()


@pytest.mark.parametrize("L,k1,sm", [(32, 7, 16), (32, 8, 24), (32, 12, 16), (32, 6, 0), (32, 32, 16), (24, 16, 24), (20, 16, 16), (16, 16, 16),
                                      (16, 4, 0), (48, 16, 16), (64, 16, 0), (40, 5, 3), (120, 16, 16)])
def test_parameter_variants(L, k1, sm):
    """other (L, k1, skipMargin) than MBGC's defaults: plain (odd k1) and Exp variants, K = 12..56, K < k1"""
    from scenarios import growing_collection, wrapping_collection
    for sc in (growing_collection(seed=31, n=4, length=60_000), wrapping_collection(seed=32, n=6, length=50_000, limit=260_001)):
        o = OracleMatcher(sc["limit"], L, k1, 1, sm)
        g = CudaFace(sc["limit"], L, k1, 1, sm)
        try:
            for m in (o, g):
                m.min_len = L
            want, sw = replay(o, sc, min_len=L)
            got, sg = replay(g, sc, min_len=L)
            assert sg == sw
            for k, (a, b) in enumerate(zip(got, want)):
                assert np.array_equal(a, b), f"op {k}: " + _diff(a, b)
            assert np.array_equal(g.table(), o.table())
        finally:
            g.close()
            o.close()


def test_batch_equals_single():
    sc = all_scenarios()[0]
    g = CudaFace(sc["limit"])
    try:
        one, _ = replay(g, sc, batch=False)
    finally:
        g.close()
    g = CudaFace(sc["limit"])
    try:
        g.disable_sliding_window()
        g.load_ref(sc["ops"][0][1], rc=True)
        texts = [op[1] for op in sc["ops"] if op[0] == "match"]
        many = g.match_batch(texts)
    finally:
        g.close()
    assert len(one) == len(many)
    for a, b in zip(one, many):
        assert np.array_equal(a, b)


def test_protein_kmer16():
    """L = 16 -> K = 12 (the reference's protein profile, MGMP_Params.h:27)"""
    rng = np.random.default_rng(77)
    alpha = np.frombuffer(b"ACDEFGHIKLMNPQRSTVWY", dtype=np.uint8)
    base = alpha[rng.integers(0, 20, 60_000)]
    t = base.copy()
    t[rng.integers(0, t.size, 600)] = ord("X")
    o = OracleMatcher(1 << 22, 16, 16, 1, 16)
    g = CudaFace(1 << 22, 16, 16, 1, 16)
    try:
        for m in (o, g):
            m.disable_sliding_window()
            m.load_ref(base, rc=False)
        assert np.array_equal(g.match(t, 16), o.match(t, 16))
        assert np.array_equal(g.table(), o.table())
    finally:
        g.close()
        o.close()


def test_min_match_length_error():
    import mbgc_b200
    g = CudaFace(1 << 22)
    try:
        with pytest.raises(mbgc_b200.MatcherError):
            g.match(np.zeros(100, np.uint8), 20)  # < K = 28: the reference exits (ref .cpp:480-483)
    finally:
        g.close()


@pytest.mark.skipif(not have_fixtures(), reason="reference fixtures not present (oracle/_ref/fixtures)")
@pytest.mark.parametrize("skip_margin", [16, 24])
def test_listeria_fixture(skip_margin):
    """config 1: the reference's shipped example (3 Listeria genomes, N runs, 16-27 contigs per file)"""
    import os
    files = [os.path.join(FIXTURES, l.strip()) for l in open(os.path.join(FIXTURES, "seqlist.txt"))]
    recs = [read_fasta(f) for f in files]
    g0 = b"".join(s for _, s in recs[0])
    limit = mgmp_ref_limit(3, len(g0))
    o = OracleMatcher(limit, 32, 16, 1, skip_margin)
    g = CudaFace(limit, 32, 16, 1, skip_margin)
    try:
        for m in (o, g):
            m.disable_sliding_window()
            m.load_ref(g0, rc=True)
        total = 0
        for fi in (1, 2):
            for _, s in recs[fi]:
                a, b = g.match(s), o.match(s)
                assert np.array_equal(a, b), _diff(a, b)
                total += len(a)
                for m in (o, g):
                    m.load_ref(s, rc=True)
            for m in (o, g):
                m.load_separator(0)
        assert total == {16: 29712, 24: 29715}[skip_margin]  # the reference's own counts (tests/golden/README.md)
        assert np.array_equal(g.table(), o.table())
    finally:
        g.close()
        o.close()


def test_synthetic_100x_near_batch():
    """config 2 shape at reduced size: 12 x 1 Mbp near-identical genomes, frozen reference, one batch"""
    gs = synth_collection(12, 1_000_000, 1234, 1e-4, 1e-5)
    limit = mgmp_ref_limit(12, gs[0].size)
    o = OracleMatcher(limit)
    g = CudaFace(limit)
    try:
        for m in (o, g):
            m.disable_sliding_window()
            m.load_ref(gs[0], rc=True)
        got = g.match_batch(gs[1:])
        for a, t in zip(got, gs[1:]):
            b = o.match(t)
            assert np.array_equal(a, b), _diff(a, b)
    finally:
        g.close()
        o.close()


def test_grown_reference_long_islands():
    """A reference that holds several similar copies of the genome (G0 + two divergent genomes, each with its reverse
    complement): seeds of the copies' diagonals interleave, every run is a single seed linked to a run a few places
    away (LINK_WINDOW), and the islands between two mismatches of a near-identical target hold thousands of runs
    (k_resolve_big).  Match lists against the oracle, one text at a time and as a batch, with and without the fast path."""
    import os
    base = synth_base(400_000, 77)
    copies = [synth_mutant(base, 7000 + i, 1e-2, 1e-4) for i in range(2)]
    targets = [synth_mutant(base, 7100 + i, p, q) for i, (p, q) in enumerate(((1e-4, 1e-5), (3e-5, 0), (1e-3, 1e-4)))]
    targets.append(synth_mutant(copies[0], 7200, 1e-4, 1e-5))   # near-identical to one of the appended copies
    limit = mgmp_ref_limit(8, base.size)
    for fast in ("0", "1"):
        os.environ["MBGC_B200_FASTPATH"] = fast
        o = OracleMatcher(limit, 32, 16, 1, 16)
        g = CudaFace(limit, 32, 16, 1, 16)
        try:
            for m in (o, g):
                m.disable_sliding_window()
                m.load_ref(base, rc=True)
                for c in copies:
                    m.load_ref(c, rc=True)
                    m.load_separator(0)
            want = [o.match(t) for t in targets]
            assert max(len(w) for w in want) > 10
            for t, w in zip(targets, want):
                got = g.match(t)
                assert np.array_equal(got, w), _diff(got, w)
            for got, w in zip(g.match_batch(targets), want):
                assert np.array_equal(got, w), _diff(got, w)
        finally:
            os.environ.pop("MBGC_B200_FASTPATH", None)
            g.close()
            o.close()


def test_insert_sequence_rebase(monkeypatch):
    """The table orders racing index samples by a 32-bit sequence number; when the numbers run out the table
    forgets them (epoch rebase) and counting restarts.  With the limit lowered to 60 000 samples, eight appended genomes
    (25 000 samples each with their reverse complements, overlapping k-mers racing for slots all the time) cross it several
    times: table, match lists and cursors must not notice."""
    monkeypatch.setenv("MBGC_B200_SEQ_LIMIT", "60000")
    base = synth_base(200_000, 88)
    muts = [synth_mutant(base, 8800 + i, 5e-3, 5e-4) for i in range(8)]
    probes = [synth_mutant(base, 8900 + i, 2e-3, 2e-4) for i in range(3)]
    limit = mgmp_ref_limit(16, base.size)
    o = OracleMatcher(limit, 32, 16, 1, 16)
    g = CudaFace(limit, 32, 16, 1, 16)
    try:
        for m in (o, g):
            m.disable_sliding_window()
            m.load_ref(base, rc=True)
        for i, t in enumerate(muts):
            for m in (o, g):
                m.load_ref(t, rc=True)
                m.load_separator(0)
            if i % 3 == 2:
                a, b = g.match(probes[i // 3]), o.match(probes[i // 3])
                assert np.array_equal(a, b), _diff(a, b)
        assert np.array_equal(g.table(), o.table()), "hash table differs after the rebases"
        for t in probes:
            a, b = g.match(t), o.match(t)
            assert np.array_equal(a, b), _diff(a, b)
    finally:
        g.close()
        o.close()
