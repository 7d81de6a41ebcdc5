"""GPU: the sharded path with its exchange step on real matchers.  Two ranks (gloo for the control
plane; both replicas live on cuda:0 here, one GPU each under torchrun) run mbgc_b200.dist.match_speculative:
optimistic batching with reference extensions broadcast between the ranks.  Result = the sequential schedule of
the CPU oracle, and both replicas end with the same index."""
import os
import socket
import sys

import numpy as np
import pytest

from test_multirank_cpu import _OracleAsMatcher, _collection, _rule

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
