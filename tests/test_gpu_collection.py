"""The collection pass (mbgcg_seq_*, row a9 of SURVEY.md 8): the reference's sequential (-t 1) target schedule run over
speculative windows on 1..N devices must produce the SAME six streams, meta data and matcher state as the
one-contig-at-a-time schedule -- here the oracle's emulation of it (tests/seq_schedule.py), which reproduces the stream
dumps of the stock `mbgc-dev v -Q -D` (tests/test_oracle_golden.py)."""
import os

import numpy as np
import pytest

import mbgc_b200
from mbgc_b200.matcher import SeqConfig
from mbgc_b200.synth import synth_collection, write_fasta
from seq_schedule import STREAMS, encode_streams, reference_factor
from util import FIXTURES, have_fixtures, read_fasta

pytestmark = pytest.mark.gpu


def n_devices():
    import torch
    return torch.cuda.device_count()


def gpu_streams(paths, mode, devices, chunk_bytes, pinned=True, max_window=0, lit_ext=None):
    """the GPU side of seq_schedule.encode_streams: same initial reference, then the pass over the file images"""
    recs0 = read_fasta(paths[0])
    c0 = recs0[0][1]
    basic = max(os.path.getsize(paths[0]), 1 << 21)
    limit = min(reference_factor(len(paths)) * basic * 2, 2**32 - 1)
    limit = max(limit, len(c0))
    g = mbgc_b200.Matcher(limit, 32, 16, 1, 24 if mode >= 2 else 16, devices=devices)
    try:
        g.disableSlidingWindow()
        g.loadRef(c0, True, True, 0)
        if lit_ext is not None:   # MBGC_Encoder::processLiteral on (row a12)
            g.setLiteralRefExt(lit_ext[0], lit_ext[1], 128)
        images = [open(p, "rb").read() for p in paths]
        out, info = g.compressSequential(images, SeqConfig.for_mode(mode, chunk_bytes, max_window), pinned=pinned)
        info["ref_length"] = g.getRefLength()
        info["loaded"] = g.getLoadedRefLength()
        info["table"] = g.table()
        return out, info
    finally:
        g.close()


def check_against_oracle(paths, mode, devices, chunk_bytes, lit_ext=None, **kw):
    want, wstats = encode_streams(paths, mode, lit_ext=lit_ext)
    got, info = gpu_streams(paths, mode, devices, chunk_bytes, lit_ext=lit_ext, **kw)
    lit0 = read_fasta(paths[0])[0][1]
    assert want["literals"][:len(lit0)] == lit0
    want = dict(want)
    want["literals"] = want["literals"][len(lit0) + 1:]  # the G0 bootstrap (contig + SEQ mark) is the encoder's
    for name in STREAMS:
        assert got[name] == want[name], f"stream {name} differs ({len(got[name])} vs {len(want[name])} bytes)"
    assert info["unmatched"] == wstats["unmatched"]
    assert info["matches"] == wstats["matches"]
    assert info["ref_length"] == wstats["ref_length"]
    recs = [read_fasta(p) for p in paths]
    assert info["seqCounts"] == [len(r) for r in recs]
    assert [h.decode() for _, h in info["headers"]] == [h for r in recs for h, _ in r]
    assert all(s == -1 for s in info["status"])
    return info


@pytest.fixture(scope="module")
def mixed_dir(tmp_path_factory):
    """divergent genomes (every contig extends the reference), then near-identical ones (nothing does), some files with
    several contigs, one with a short last line and one tiny contig"""
    d = tmp_path_factory.mktemp("coll")
    gs = synth_collection(7, 300_000, 4321, 1e-2, 1e-4)
    gs += synth_collection(9, 300_000, 4321, 1e-4, 1e-5, first=40)[1:]
    paths = []
    for i, g in enumerate(gs):
        p = os.path.join(d, f"g{i:04d}.fna")
        if i % 4 == 1:  # three contigs
            a, b = g.size // 3, 2 * g.size // 3
            with open(p, "wb") as f:
                for k, part in enumerate((g[:a], g[a:b], g[b:])):
                    tmp = os.path.join(d, "part.tmp")
                    write_fasta(tmp, f"syn_{i:04d}.{k} contig {k}", part, width=70)
                    f.write(open(tmp, "rb").read())
        elif i % 4 == 2:
            with open(p, "wb") as f:
                tmp = os.path.join(d, "part.tmp")
                write_fasta(tmp, f"syn_{i:04d} main", g, width=80)
                f.write(open(tmp, "rb").read())
                f.write(b">tiny\nACGTACGTAC\n")
        else:
            write_fasta(p, f"syn_{i:04d} synthetic E. coli-like genome {i}", g)
        paths.append(p)
    return paths


@pytest.mark.parametrize("mode", [0, 1, 3])
def test_streams_equal_sequential_schedule(mixed_dir, mode):
    info = check_against_oracle(mixed_dir, mode, [0], 1 << 20)
    st = info["stats"]
    assert st["extensions"] > 0 and st["windows"] > 1 and st["launches"] > 0
    assert st["pageableBytes"] == 0, "images from mbgcg_image_alloc must travel by DMA from pinned memory"


def test_one_big_window_and_pageable_images(mixed_dir):
    info = check_against_oracle(mixed_dir, 1, [0], 1 << 30, pinned=False)
    assert info["stats"]["windows"] == 1 and info["stats"]["pageableBytes"] > 0


def test_small_speculation_window(mixed_dir):
    check_against_oracle(mixed_dir, 1, [0], 2 << 20, max_window=2)


@pytest.mark.skipif(not have_fixtures(), reason="reference fixtures not present")
@pytest.mark.parametrize("mode", [1, 2])
def test_listeria_streams(mode):
    paths = [os.path.join(FIXTURES, l.strip()) for l in open(os.path.join(FIXTURES, "seqlist.txt"))]
    check_against_oracle(paths, mode, [0], 2 << 20)


def test_two_replicas_same_streams(mixed_dir):
    if n_devices() < 2:
        pytest.skip("needs two CUDA devices")
    info = check_against_oracle(mixed_dir, 1, [0, 1], 1 << 20)
    assert info["stats"]["extensions"] > 0


def test_replicas_hold_identical_tables(mixed_dir):
    """after a pass with extensions every replica's index equals the single-device one (ordered inserts everywhere)"""
    if n_devices() < 2:
        pytest.skip("needs two CUDA devices")
    _, one = gpu_streams(mixed_dir[:6], 1, [0], 1 << 20)
    _, two = gpu_streams(mixed_dir[:6], 1, [1, 0], 1 << 20)  # replica 0 on device 1
    assert np.array_equal(one["table"], two["table"])
    assert one["loaded"] == two["loaded"]


def test_bad_files_report_the_readers_status(tmp_path):
    good = tmp_path / "a.fna"
    write_fasta(str(good), "a", synth_collection(1, 50_000, 9, 0, 0)[0])
    g = mbgc_b200.Matcher(1 << 23, 32, 16, 1, 16, device=0)
    try:
        g.disableSlidingWindow()
        g.loadRef(read_fasta(str(good))[0][1], True, True, 0)
        images = [open(good, "rb").read(), b"", b"not fasta\n", b">x\nACGT\n\nAC\n"]
        _, info = g.compressSequential(images, SeqConfig.for_mode(1, 1 << 20))
        assert info["status"] == [-1, -1, -3, -4]
        assert info["seqCounts"][:2] == [1, 0]
    finally:
        g.close()


def _files_pass(paths, mode, pack, chunk_bytes=1 << 20):
    recs0 = read_fasta(paths[0])
    c0 = recs0[0][1]
    limit = max(min(reference_factor(len(paths)) * max(os.path.getsize(paths[0]), 1 << 21) * 2, 2**32 - 1), len(c0))
    g = mbgc_b200.Matcher(limit, 32, 16, 1, 24 if mode >= 2 else 16, device=0)
    try:
        g.disableSlidingWindow()
        g.loadRef(c0, True, True, 0)
        out, info = g.compressFiles(paths, SeqConfig.for_mode(mode, chunk_bytes), readers=4, pack=pack)
        info["table"] = g.table()
        return out, info
    finally:
        g.close()


@pytest.mark.parametrize("mode", [1, 3])
def test_packed_file_images_equal_plain_ones(mixed_dir, tmp_path, mode):
    """file images that cross PCIe 4 bases per byte (mbgcg_image_pack from an mmap of the file, line ends as one run,
    rebuilt byte for byte on the device) give the same streams, headers, meta data and index as images read as they are --
    including files the packer refuses (CRLF line ends, lower case: too many exceptional bytes), N runs across line ends,
    a file without a final line end, an empty file and one that is not FASTA"""
    g = synth_collection(2, 120_000, 77, 2e-3, 2e-4)[1]
    extra = []
    withN = g.copy()
    withN[1000:9000] = ord("N")
    withN[50_000] = ord("R")
    p = tmp_path / "withN.fna"
    write_fasta(str(p), "genome with an N run and an IUPAC code", withN, width=60)
    extra.append(str(p))
    p = tmp_path / "nofinal.fna"
    write_fasta(str(p), "no final line end", g[:70_001], width=80)
    open(p, "wb").write(open(p, "rb").read().rstrip(b"\n"))
    extra.append(str(p))
    p = tmp_path / "crlf.fna"
    write_fasta(str(p), "crlf", g[:40_000], width=80)
    open(p, "wb").write(open(p, "rb").read().replace(b"\n", b"\r\n"))
    extra.append(str(p))
    p = tmp_path / "lower.fna"
    write_fasta(str(p), "soft-masked", g[:30_000], width=80)
    raw = open(p, "rb").read()
    open(p, "wb").write(raw[:20] + raw[20:].lower())
    extra.append(str(p))
    p = tmp_path / "empty.fna"
    open(p, "wb").close()
    extra.append(str(p))
    p = tmp_path / "notfasta.txt"
    open(p, "wb").write(b"hello\nworld\n" * 100)
    extra.append(str(p))
    p = tmp_path / "oneline.fna"
    open(p, "wb").write(b">one line\n" + bytes(g[:5000]) + b"\n>second\n" + bytes(g[5000:5100]) + b"\n")
    extra.append(str(p))
    paths = list(mixed_dir[:9]) + extra
    plain, pi = _files_pass(paths, mode, False)
    packed, qi = _files_pass(paths, mode, True)
    for name in STREAMS:
        assert packed[name] == plain[name], f"stream {name} differs between packed and plain images"
    for k in ("seqCounts", "dnaLineLens", "status", "headers", "refExtSizes", "unmatched", "matches", "extensions", "digest"):
        assert qi[k] == pi[k], k
    assert np.array_equal(qi["table"], pi["table"])
    assert qi["stats"]["wireBytes"] < 0.45 * pi["stats"]["wireBytes"], "most of the bytes must travel packed"
    assert pi["stats"]["wireBytes"] == pi["stats"]["imageBytes"]


def test_files_pass_two_replicas(mixed_dir):
    """the library's own readers feeding two replicas of one process (packed images, extensions device to device)"""
    if n_devices() < 2:
        pytest.skip("needs two CUDA devices")
    c0 = read_fasta(mixed_dir[0])[0][1]
    limit = max(min(reference_factor(len(mixed_dir)) * max(os.path.getsize(mixed_dir[0]), 1 << 21) * 2, 2**32 - 1), len(c0))
    res = []
    for devices in ([0], [0, 1]):
        g = mbgc_b200.Matcher(limit, 32, 16, 1, 16, devices=devices)
        try:
            g.disableSlidingWindow()
            g.loadRef(c0, True, True, 0)
            out, info = g.compressFiles(list(mixed_dir), SeqConfig.for_mode(1, 1 << 20), readers=6, pack=True)
            res.append((out, info, g.table()))
        finally:
            g.close()
    for name in STREAMS:
        assert res[0][0][name] == res[1][0][name], name
    assert res[0][1]["digest"] == res[1][1]["digest"] and res[0][1]["extensions"] == res[1][1]["extensions"] > 0
    assert np.array_equal(res[0][2], res[1][2])


def test_files_pass_errors_and_empty_list(mixed_dir, tmp_path):
    """mbgcg_seq_run_files: a file that cannot be read ends the pass with an error (and leaves the handle usable), an empty
    list is an empty pass"""
    c0 = read_fasta(mixed_dir[0])[0][1]
    g = mbgc_b200.Matcher(1 << 24, 32, 16, 1, 16, device=0)
    try:
        g.disableSlidingWindow()
        g.loadRef(c0, True, True, 0)
        with pytest.raises(mbgc_b200.MatcherError):
            g.compressFiles([mixed_dir[8], str(tmp_path / "does_not_exist.fna"), mixed_dir[9]], SeqConfig.for_mode(1, 1 << 20), readers=3)
        out, info = g.compressFiles([], SeqConfig.for_mode(1, 1 << 20))
        assert info["seqCounts"] == [] and all(v == b"" for v in out.values())
        out, info = g.compressFiles(list(mixed_dir[8:11]), SeqConfig.for_mode(1, 1 << 20), readers=2)
        assert info["seqCounts"] == [len(read_fasta(p)) for p in mixed_dir[8:11]] and info["status"] == [-1, -1, -1]
    finally:
        g.close()


# ---- the batch-frozen schedule of the reference's parallel mode ---------------------------------------------------------

def check_parallel_schedule(paths, mode, devices, chunk_bytes, outrun):
    from par_schedule import encode_streams_parallel
    from mbgc_b200.matcher import EV_FILE_END
    recs0 = read_fasta(paths[0])
    basic = max(os.path.getsize(paths[0]), 1 << 21)
    limit = max(min(reference_factor(len(paths)) * basic * 2, 2**32 - 1), len(recs0[0][1]))
    g = mbgc_b200.Matcher(limit, 32, 16, 1, 24 if mode >= 2 else 16, devices=devices)
    try:
        g.disableSlidingWindow()
        g.loadRef(recs0[0][1], True, True, 0)
        cfg = SeqConfig.for_mode(mode, chunk_bytes, parallel=True, outrun=outrun)
        windows = {}

        def on_event(ev):
            if ev.type == EV_FILE_END:
                windows.setdefault(int(ev.window), []).append(int(ev.file))

        got, info = g.compressSequential([open(p, "rb").read() for p in paths], cfg, on_event=on_event)
        ref_length = g.getRefLength()
    finally:
        g.close()
    wins = [windows[k] for k in sorted(windows)]
    want, st = encode_streams_parallel(paths, mode, limit, wins, cfg.allowedTargetsOutrunForDissimilarContigs,
                                       cfg.unmatchedFractionFactorTweakForDissimilarContigs)
    for name in STREAMS:
        assert got[name] == want[name], f"stream {name} differs ({len(got[name])} vs {len(want[name])} bytes)"
    assert info["unmatched"] == st["unmatched"] and info["matches"] == st["matches"]
    assert info["lockPos"] == st["locks"] and info["refExtSizes"] == st["ref_ext"]
    assert ref_length == st["ref_length"]
    return info, wins, st


@pytest.mark.parametrize("mode", [0, 1, 2])
def test_parallel_schedule_streams(mixed_dir, mode):
    info, wins, st = check_parallel_schedule(mixed_dir, mode, [0], 1 << 20, outrun=4)
    assert len(wins) > 2 and max(len(w) for w in wins) > 1 and info["extensions"] > 0


def test_parallel_schedule_two_replicas(mixed_dir):
    if n_devices() < 2:
        pytest.skip("needs two CUDA devices")
    check_parallel_schedule(mixed_dir, 1, [0, 1], 1 << 20, outrun=6)
