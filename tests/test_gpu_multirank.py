"""GPU: the sharded path with its exchange step on real matchers.  Two ranks (gloo for the control
plane; both replicas live on cuda:0 here, one GPU each under torchrun) run mbgc_b200.dist.match_speculative:
optimistic batching with reference extensions broadcast between the ranks.  Result = the sequential schedule of
the CPU oracle, and both replicas end with the same index."""
import os
import socket
import sys

import numpy as np
import pytest

from test_multirank_cpu import _OracleAsMatcher, _collection, _compress_files, _rule, _write_files

pytestmark = pytest.mark.gpu


def _worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
    import torch.distributed as dist
    import mbgc_b200
    from mbgc_b200.dist import match_speculative
    dist.init_process_group("gloo", rank=rank, world_size=world)
    base, texts = _collection()
    m = mbgc_b200.Matcher(1 << 22, 32, 16, 1, 16, device=0)
    m.disableSlidingWindow()
    m.loadRef(base, True, True, 0)
    lists, rounds = match_speculative(m, texts, _rule, window=4)
    tsum = int(m.table().astype(np.uint64).sum())
    pos = m.getLoadingPosition()
    dist.barrier()
    q.put((rank, [a.tolist() for a in lists], rounds, tsum, pos))
    m.close()
    dist.destroy_process_group()


def test_speculative_batching_two_ranks_gpu():
    import torch.multiprocessing as mp
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=600) for _ in range(2))
    for p in procs:
        p.join(timeout=120)
        assert p.exitcode == 0
    base, texts = _collection()
    o = _OracleAsMatcher(1 << 22)
    o.loadRef(base, True)
    want = []
    for t in texts:
        a = o.o.match(t)
        want.append(a)
        ext, rc = _rule(len(t), a)
        if ext:
            o.loadRef(t, rc)
    for rank, lists, rounds, tsum, pos in res:
        for a, b in zip(lists, want):
            assert np.array_equal(np.array(a, dtype=np.uint64).reshape(-1, 3), b)
        assert tsum == int(o.o.table().astype(np.uint64).sum()) and pos == o.o.loading_position
        assert rounds < len(texts)


def _compress_worker(rank, world, port, q, limit):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
    import torch.distributed as dist
    import mbgc_b200
    from mbgc_b200.dist import compress_speculative
    from mbgc_b200.matcher import PackParams
    dist.init_process_group("gloo", rank=rank, world_size=world)
    files = _compress_files()
    m = mbgc_b200.Matcher(limit, 32, 16, 1, 16, device=0)
    m.disableSlidingWindow()
    m.loadRef(files[0][0], True, True, 0)
    streams, stats = compress_speculative(m, files, PackParams.for_mode(1), window=5)
    pos = m.getLoadingPosition()
    dist.barrier()
    q.put((rank, streams, stats, pos))
    m.close()
    dist.destroy_process_group()


def test_sharded_compress_two_ranks_gpu(tmp_path):
    """both stages of the hot path (match + K6 pack) on real matchers, contigs sharded over two ranks, reference
    extensions by the reference's rule on the packer's unmatched count: the six streams of the sequential schedule"""
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
    res = sorted((q.get(timeout=600) for _ in range(2)), key=lambda x: x[0])
    for p in procs:
        p.join(timeout=120)
        assert p.exitcode == 0
    boot = bytes(files[0][0]) + bytes([SEQ_SEPARATOR_MARK])
    for rank, streams, stats, pos in res:
        assert 0 < stats["extensions"] < stats["contigs"] and stats["rounds"] < stats["contigs"]
        for name, data in want.items():
            got = (boot if name == "literals" else b"") + streams[name]
            assert got == data, f"rank {rank} stream {name} differs ({len(got)} vs {len(data)} bytes)"


def _shard_worker(rank, world, port, q, limit, paths):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
    import torch.distributed as dist
    import mbgc_b200
    from mbgc_b200.dist import compress_sharded_files
    from mbgc_b200.matcher import SeqConfig
    from util import read_fasta
    dist.init_process_group("gloo", rank=rank, world_size=world)
    m = mbgc_b200.Matcher(limit, 32, 16, 1, 16, device=0)
    m.disableSlidingWindow()
    m.loadRef(read_fasta(paths[0])[0][1], True, True, 0)
    streams, info, ext = compress_sharded_files(m, paths, SeqConfig.for_mode(1, 1 << 20), readers=3)
    dist.barrier()
    q.put((rank, streams, info["unmatched"], info["matches"], ext))
    m.close()
    dist.destroy_process_group()


def test_sharded_files_two_ranks_gpu(tmp_path):
    """the collection pass with every rank reading only its own files (near-identical collection: nothing extends the
    reference): the per-rank streams concatenated in rank order are the streams of the sequential schedule"""
    import torch.multiprocessing as mp
    from mbgc_b200.synth import synth_collection, write_fasta
    from seq_schedule import STREAMS, encode_streams
    from util import read_fasta
    gs = synth_collection(9, 200_000, 99, 1e-4, 1e-5)
    paths = []
    for i, g in enumerate(gs):
        p = os.path.join(str(tmp_path), f"g{i:04d}.fna")
        write_fasta(p, f"syn_{i:04d} near-identical genome {i}", g)
        paths.append(p)
    want, wstats = encode_streams(paths, 1)
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_shard_worker, args=(r, 2, port, q, wstats["limit"], paths)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted((q.get(timeout=600) for _ in range(2)), key=lambda x: x[0])
    for p in procs:
        p.join(timeout=120)
        assert p.exitcode == 0
    lit0 = read_fasta(paths[0])[0][1]
    for name in STREAMS:
        got = b"".join(r[1][name] for r in res)
        exp = want[name][len(lit0) + 1:] if name == "literals" else want[name]
        assert got == exp, f"stream {name} differs ({len(got)} vs {len(exp)} bytes)"
    assert sum(r[2] for r in res) == wstats["unmatched"] and sum(r[3] for r in res) == wstats["matches"]
    assert res[0][4] == [0, 0]
