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
