"""N > 1 path on CPU: two gloo ranks shard the targets of one batch, match them (the CPU oracle stands in for
the GPU matcher: same frozen-state contract) and exchange counts / lists / timings exactly as the GPU run does
with nccl.  Result must equal the single-process run in target order."""
import os
import socket
import sys

import numpy as np
import pytest

from mbgc_b200.dist import shard_range


def test_shard_range_partitions():
    for n in (0, 1, 7, 8, 100, 1000):
        for w in (1, 2, 3, 8):
            spans = [shard_range(n, r, w) for r in range(w)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
            sizes = [e - b for b, e in spans]
            assert max(sizes) - min(sizes) <= 1


def _worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
    import torch.distributed as dist
    from mbgc_b200.dist import gather_in_target_order, max_over_ranks, sum_over_ranks
    from mbgc_b200.synth import synth_collection
    from util import OracleMatcher
    dist.init_process_group("gloo", rank=rank, world_size=world)
    gs = synth_collection(8, 120_000, 321, 2e-3, 2e-4)
    m = OracleMatcher(1 << 22)
    m.disable_sliding_window()
    m.load_ref(gs[0], rc=True)  # replicated index: every rank builds the same frozen state
    targets = gs[1:]
    b, e = shard_range(len(targets), rank, world)
    local = [m.match(t) for t in targets[b:e]]
    allm = gather_in_target_order(local, len(targets))
    total = sum_over_ranks(sum(len(a) for a in local))
    slow = max_over_ranks(10.0 + rank)
    dist.barrier()
    if rank == 0:
        q.put(([a.tolist() for a in allm], total, slow))
    dist.destroy_process_group()


def test_two_ranks_equal_single_process():
    import torch.multiprocessing as mp
    from mbgc_b200.synth import synth_collection
    from util import OracleMatcher
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    got, total, slow = q.get(timeout=300)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    gs = synth_collection(8, 120_000, 321, 2e-3, 2e-4)
    m = OracleMatcher(1 << 22)
    m.disable_sliding_window()
    m.load_ref(gs[0], rc=True)
    want = [m.match(t) for t in gs[1:]]
    assert len(got) == len(want) and total == sum(len(a) for a in want) and slow == 11.0
    for a, b in zip(got, want):
        assert np.array_equal(np.array(a, dtype=np.uint64).reshape(-1, 3), b)


class _OracleAsMatcher:
    """the CPU oracle behind the two Matcher methods match_speculative uses (tests only)"""

    def __init__(self, limit):
        from util import OracleMatcher
        self.o = OracleMatcher(limit)
        self.o.disable_sliding_window()

    def loadRef(self, text, rc, add_sep=True, sep=0):
        self.o.load_ref(np.asarray(text, dtype=np.uint8), rc=bool(rc))

    def matchBatch(self, texts, min_len=32):
        return [self.o.match(t, min_len) for t in texts]


def _rule(n, m):
    """stand-in for isContigProperForRefExtension (MGMP_Params.h:179-190): unmatched * 128 > length"""
    matched = int(m[:, 1].sum()) if len(m) else 0
    unmatched = n - matched
    return unmatched * 128 > n, unmatched * 8 > n


def _collection():
    from mbgc_b200.synth import synth_base, synth_mutant
    base = synth_base(80_000, 77)
    other = synth_base(60_000, 78)
    texts = []
    for i in range(10):
        src = base if i % 4 != 2 else other                       # every 4th text is novel: it extends the reference
        texts.append(synth_mutant(src, 500 + i, 2e-4 if i % 3 else 2e-2, 1e-5))
    return base, texts


def _spec_worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
    import torch.distributed as dist
    from mbgc_b200.dist import match_speculative
    dist.init_process_group("gloo", rank=rank, world_size=world)
    base, texts = _collection()
    m = _OracleAsMatcher(1 << 22)
    m.loadRef(base, True)
    lists, rounds = match_speculative(m, texts, _rule, window=4)
    table = m.o.table()
    dist.barrier()
    q.put((rank, [a.tolist() for a in lists], rounds, int(table.astype(np.uint64).sum()), m.o.loading_position))
    dist.destroy_process_group()


def test_speculative_batching_equals_sequential_schedule():
    """optimistic batching over 2 ranks with reference extensions broadcast between them gives the
    match lists of the one-text-at-a-time schedule, and both replicas end with the same index"""
    import torch.multiprocessing as mp
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_spec_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=300) for _ in range(2))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    base, texts = _collection()
    m = _OracleAsMatcher(1 << 22)
    m.loadRef(base, True)
    want, extensions = [], 0
    for t in texts:
        a = m.o.match(t)
        want.append(a)
        ext, rc = _rule(len(t), a)
        if ext:
            m.loadRef(t, rc)
            extensions += 1
    assert 0 < extensions < len(texts)
    for rank, lists, rounds, tsum, pos in res:
        assert len(lists) == len(want)
        for a, b in zip(lists, want):
            assert np.array_equal(np.array(a, dtype=np.uint64).reshape(-1, 3), b)
        assert tsum == int(m.o.table().astype(np.uint64).sum()) and pos == m.o.loading_position
        assert rounds < len(texts)  # some windows committed more than one text


# ------------------------------------------------------------------------------------------------------
# the whole hot path sharded: match + pack + the reference's extension rule (mbgc_b200.dist.compress_speculative)
# ------------------------------------------------------------------------------------------------------

class _OracleAsCompressor(_OracleAsMatcher):
    """the CPU oracle behind the Matcher methods compress_speculative uses (tests only): match lists stay inside the
    adapter between matchBatch(fetch=False) and packBatch(matches=None), like they stay on the device"""

    def getLoadingPosition(self):
        return self.o.loading_position

    def getLoadedRefLength(self):
        return self.o.loaded_ref_length

    def loadSeparator(self, sep=0):
        self.o.load_separator(sep)

    def matchBatch(self, texts, min_len=32, fetch=True):
        self._lists = [self.o.match(t, min_len) for t in texts]
        return self._lists if fetch else np.array([len(a) for a in self._lists], dtype=np.uint64)

    def packBatch(self, pd, texts, matches=None, lockPos=None, loadedPosArr=(), lens=None):
        import ctypes as C
        from util import PackParams, oracle_lib, u8p
        assert matches is None
        lib = oracle_lib()
        po = PackParams(pd.enableExtensionsWithMismatches, pd.gapBreakingMatchMinLength, pd.gapDepthOffsetEncoding,
                        pd.gapDepthMismatchesEncoding, pd.lazyDecompressionSupport, pd.frugal64bitLenEncoding,
                        pd.enable40bitReference, pd.mismatchesWithExclusion, pd.mmsMismatchesInitialScore,
                        pd.mmsMismatchesScoreThreshold, pd.mmsMismatchPenalty, pd.mmsMatchBonus, 32, 128)
        names = ("literals", "mapOff", "mapOff5", "mapLen", "gapDelta", "gapFlags")
        out = {n: [] for n in names}
        out["unmatched"], out["kept"] = [], []
        arr = (C.c_uint64 * max(1, len(loadedPosArr)))(*loadedPosArr)
        for t, mt in zip(texts, self._lists):
            st = lib.mbo_streams_create()
            a = np.ascontiguousarray(t, dtype=np.uint8)
            buf = np.ascontiguousarray(mt, dtype=np.uint64).copy() if len(mt) else np.zeros((1, 3), dtype=np.uint64)
            n = C.c_uint64(len(mt))
            unmatched = lib.mbo_process_matches(st, C.byref(po), self.o.h, buf.ctypes.data_as(C.c_void_p), C.byref(n),
                                                a.ctypes.data_as(u8p), a.size, 2**64 - 1, arr, len(loadedPosArr))
            for i, name in enumerate(names):
                k = lib.mbo_stream_size(st, i)
                ptr = lib.mbo_stream_data(st, i)
                out[name].append(bytes(np.ctypeslib.as_array(C.cast(ptr, u8p), shape=(k,))) if k else b"")
            out["unmatched"].append(unmatched)
            out["kept"].append(n.value)
            lib.mbo_streams_destroy(st)
        return out


def _compress_files():
    """6 files x 1..3 contigs: near-identical ones (never extend), a divergent and two novel ones (extend, with and
    without the reverse complement), a tiny contig"""
    from mbgc_b200.synth import synth_base, synth_mutant
    base = synth_base(70_000, 177)
    other = synth_base(40_000, 178)
    files = [[base],
             [synth_mutant(base, 601, 2e-4, 1e-5), synth_mutant(other, 602, 1e-4, 1e-5)],
             [synth_mutant(base, 603, 3e-4, 2e-5)],
             [synth_mutant(base, 604, 3e-2, 1e-3), synth_mutant(base, 605, 1e-4, 1e-5)[:31], synth_mutant(other, 606, 2e-4, 1e-5)],
             [synth_mutant(base, 607, 1e-4, 1e-5), synth_mutant(base, 608, 2e-4, 0)],
             [synth_mutant(synth_base(30_000, 179), 609, 0, 0), synth_mutant(base, 610, 5e-4, 1e-5)]]
    return files


def _write_files(d, files):
    from util import write_fasta
    paths = []
    for i, f in enumerate(files):
        path = os.path.join(d, f"f{i:03d}.fna")
        with open(path, "wb") as fh:
            for j, c in enumerate(f):
                fh.write(f">file{i} contig{j}\n".encode())
                seq = bytes(c)
                for k in range(0, len(seq), 80):
                    fh.write(seq[k:k + 80] + b"\n")
        paths.append(path)
    return paths


def _compress_worker(rank, world, port, q, limit):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
    import torch.distributed as dist
    from mbgc_b200.dist import compress_speculative
    from mbgc_b200.matcher import PackParams
    dist.init_process_group("gloo", rank=rank, world_size=world)
    files = _compress_files()
    m = _OracleAsCompressor(limit)
    m.loadRef(files[0][0], True)
    streams, stats = compress_speculative(m, files, PackParams.for_mode(1), window=5)
    dist.barrier()
    q.put((rank, streams, stats, m.o.loading_position))
    dist.destroy_process_group()


def test_sharded_compress_equals_sequential_streams(tmp_path):
    """match + pack sharded over 2 ranks with speculative windows across file boundaries == the six streams of the
    sequential (-t 1) schedule (tests/seq_schedule.py, pinned against the reference's stream dumps)"""
    import torch.multiprocessing as mp
    from seq_schedule import SEQ_SEPARATOR_MARK, encode_streams
    files = _compress_files()
    paths = _write_files(str(tmp_path), files)
    want, wstats = encode_streams(paths, 1)
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_compress_worker, args=(r, 2, port, q, wstats["limit"])) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted((q.get(timeout=300) for _ in range(2)), key=lambda x: x[0])
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    boot = bytes(files[0][0]) + bytes([SEQ_SEPARATOR_MARK])
    for rank, streams, stats, pos in res:
        assert 0 < stats["extensions"] < stats["contigs"] and stats["rounds"] < stats["contigs"]
        assert stats["unmatched"] == wstats["unmatched"]
        for name, data in want.items():
            got = (boot if name == "literals" else b"") + streams[name]
            assert got == data, f"rank {rank} stream {name} differs ({len(got)} vs {len(data)} bytes)"
