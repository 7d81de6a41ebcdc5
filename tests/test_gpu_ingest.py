"""GPU parity of the FASTA ingest (mbgcg_ingest_fasta) against the oracle's restatement of kseq_read_lossless_fasta
(oracle/mbgc_oracle.c: mbo_fasta_parse, pinned against the reference's own kseq.h in tests/test_fasta_oracle.py):
records, header bytes, sequences, DNA line length, final status -- and the ingested texts feed the match stage."""
import numpy as np
import pytest

from fasta_cases import fasta_cases
from test_fasta_oracle import oracle_parse
from util import OracleMatcher, synth_base, synth_mutant

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def matcher():
    import mbgc_b200
    m = mbgc_b200.Matcher(1 << 22, 32, 16, 1, 16)
    yield m
    m.close()


@pytest.mark.parametrize("name", sorted(fasta_cases()))
def test_ingest_equals_oracle(matcher, name):
    data = fasta_cases()[name]
    want, wdna, wmax, wstatus = oracle_parse(data)
    got = matcher.ingestFasta(data)
    assert got["status"] == wstatus, f"status {got['status']} vs oracle {wstatus}"
    assert len(got["records"]) == len(want)
    assert got["dnaLineLen"] == wdna
    for k, ((h_off, h_len, t_off, s_len), (whdr, wseq)) in enumerate(zip(got["records"], want)):
        assert data[int(h_off):int(h_off + h_len)] == whdr, f"record {k}: header differs"
        assert int(s_len) == len(wseq), f"record {k}: length {s_len} vs {len(wseq)}"
        assert bytes(matcher.copyIngested(k, s_len)) == wseq, f"record {k}: sequence differs"
        assert int(t_off) % 16 == 0


def test_ingested_texts_feed_the_match_stage():
    """a multi-record FASTA file -> device texts -> mbgcg_match_batch_device, against the oracle matching the same contigs"""
    import mbgc_b200
    base = synth_base(120_000, 55)
    contigs = [synth_mutant(base, 900 + i, 2e-3, 2e-4)[:n] for i, n in enumerate((120_000, 50_017, 80, 31, 99_999))]
    parts = []
    for i, c in enumerate(contigs):
        parts.append(b">contig%d\n" % i)
        s = bytes(c)
        parts += [s[k:k + 70] + b"\n" for k in range(0, len(s), 70)]
    data = b"".join(parts)
    o = OracleMatcher(1 << 22)
    o.disable_sliding_window()
    o.load_ref(base, rc=True)
    g = mbgc_b200.Matcher(1 << 22, 32, 16, 1, 16)
    try:
        g.disableSlidingWindow()
        g.loadRef(base, True, True, 0)
        ing = g.ingestFasta(data)
        assert ing["status"] == -1 and ing["dnaLineLen"] == 70 and len(ing["records"]) == len(contigs)
        recs = ing["records"]
        lists = g.matchBatchDevice(ing["dTexts"], recs[:, 2].copy(), recs[:, 3].copy(), 32)
        for c, got in zip(contigs, lists):
            assert np.array_equal(got, o.match(c))
        # ... and, with the lists left on the device, the packer: file image -> six streams without the texts or the
        # lists ever visiting the host
        from test_gpu_pack import STREAMS, _oracle_pack, _params
        po, pd = _params()
        loaded = [o.loading_position]
        g.matchBatchDevice(ing["dTexts"], recs[:, 2].copy(), recs[:, 3].copy(), 32, fetch=False)
        out = g.packBatch(pd, None, None, loadedPosArr=loaded, lens=recs[:, 3].copy())
        for i, c in enumerate(contigs):
            ws, wk, wu = _oracle_pack(o, po, c, o.match(c), loaded=loaded)
            for name in STREAMS:
                assert out[name][i] == ws[name], f"record {i} stream {name}"
            assert int(out["unmatched"][i]) == wu and int(out["kept"][i]) == len(wk)
    finally:
        g.close()
        o.close()


def test_big_file_roundtrip(matcher):
    """a 40 MB image with 8 records: every sequence byte-exact (the gather is an index transformation)"""
    rng = np.random.default_rng(5)
    acgt = np.frombuffer(b"ACGT", dtype=np.uint8)
    seqs = [bytes(acgt[rng.integers(0, 4, n)]) for n in (5_000_000, 4_999_999, 5_000_001, 3, 7_000_000, 80, 9_000_000, 8_999_920)]
    parts = []
    for i, s in enumerate(seqs):
        parts.append(b">syn_%04d synthetic genome %d\n" % (i, i))
        a = np.frombuffer(s, dtype=np.uint8)
        full = len(s) // 80 * 80
        lines = np.concatenate([a[:full].reshape(-1, 80), np.full((full // 80, 1), 10, dtype=np.uint8)], axis=1).tobytes()
        parts.append(lines)
        if len(s) > full:
            parts.append(s[full:] + b"\n")
    data = b"".join(parts)
    got = matcher.ingestFasta(data)
    assert got["status"] == -1 and got["dnaLineLen"] == 80 and len(got["records"]) == len(seqs)
    for k, s in enumerate(seqs):
        assert int(got["records"][k][3]) == len(s)
        assert bytes(matcher.copyIngested(k, len(s))) == s


def test_unwrapped_records(matcher):
    """single-line (unwrapped) sequences of a few MB and a very long header: the record scan is a CTA-wide search"""
    rng = np.random.default_rng(6)
    acgt = np.frombuffer(b"ACGT", dtype=np.uint8)
    seqs = [bytes(acgt[rng.integers(0, 4, n)]) for n in (3_000_001, 17, 2_500_000)]
    hdrs = [b"h" * 70_000, b"short", b""]
    data = b"".join(b">" + h + b"\n" + s + b"\n" for h, s in zip(hdrs, seqs))[:-1]   # no trailing newline
    want, wdna, _, wstatus = oracle_parse(data)
    got = matcher.ingestFasta(data)
    assert (got["status"], got["dnaLineLen"], len(got["records"])) == (wstatus, wdna, len(want)) == (-1, 0, 3)
    for k, ((h_off, h_len, _, s_len), (whdr, wseq)) in enumerate(zip(got["records"], want)):
        assert data[int(h_off):int(h_off + h_len)] == whdr and bytes(matcher.copyIngested(k, s_len)) == wseq
