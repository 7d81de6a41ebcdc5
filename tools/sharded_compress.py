#!/usr/bin/env python
"""The whole hot path sharded over the GPUs of one box: mbgc_b200.dist.compress_speculative (match stage + stream packing
+ the reference's extension rule; contigs of a speculative window sharded over the ranks, extensions broadcast) under
torchrun with the NCCL backend.

    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P \
        tools/sharded_compress.py [--genomes 64] [--length 5000000] [--tier near|div|mixed] [--window 64] [--check]

Every rank holds the same collection (as every rank of the reference-facing driver would read the same file list) and a
replica of the index.  Rank 0 prints one JSON line: input GB/s of the schedule (device time, max over ranks), rounds,
extensions, stream sizes, and -- with --check -- whether the six streams equal those of the same schedule run by rank 0
alone with window 1 (the sequential order of the reference's -t 1 encoder)."""
import argparse
import hashlib
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from mbgc_b200.synth import fasta_bytes, mgmp_ref_limit, synth_base, synth_mutant  # noqa: E402

TIERS = {"near": (1e-4, 1e-5), "div": (1e-2, 1e-4)}


def collection(args):
    base = synth_base(args.length, 1234)
    files = [[base]]
    novel = synth_base(args.length // 4, 4321)
    for i in range(1, args.genomes):
        if args.tier == "mixed":
            if i % 16 == 5:      # a novel contig: extends the reference
                files.append([synth_mutant(base, 1234 * 1000003 + i, 1e-4, 1e-5), synth_mutant(novel, 77 + i, 1e-3, 1e-4)])
                continue
            p = TIERS["div"] if i % 7 == 3 else TIERS["near"]
        else:
            p = TIERS[args.tier]
        files.append([synth_mutant(base, 1234 * 1000003 + i, *p)])
    return files


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--genomes", type=int, default=64)
    ap.add_argument("--length", type=int, default=5_000_000)
    ap.add_argument("--tier", default="near", choices=["near", "div", "mixed"])
    ap.add_argument("--window", type=int, default=64)
    ap.add_argument("--mode", type=int, default=1)
    ap.add_argument("--check", action="store_true")
    args = ap.parse_args()
    import torch
    import torch.distributed as dist
    import mbgc_b200
    from mbgc_b200.dist import compress_speculative, max_over_ranks
    from mbgc_b200.matcher import PackParams

    rank, world = int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    files = collection(args)
    nbytes = sum(fasta_bytes(c.size, 40) for f in files for c in f)
    limit = mgmp_ref_limit(args.genomes + 1, args.length)
    pd = PackParams.for_mode(args.mode)

    def run(window, sharded):
        m = mbgc_b200.Matcher(limit, 32, 16, 1, 24 if args.mode >= 2 else 16, device=local)
        m.disableSlidingWindow()
        m.loadRef(files[0][0], True, True, 0)
        spent = {}
        for name in ("matchBatch", "packBatch", "loadRef", "loadSeparator"):   # wall time per matcher entry point
            def wrap(fn=getattr(m, name), name=name):
                def timed(*a, **k):
                    t = time.perf_counter()
                    r = fn(*a, **k)
                    spent[name] = spent.get(name, 0.0) + time.perf_counter() - t
                    if os.environ.get("SHARDED_TRACE") and name == "matchBatch":
                        print(f"[rank {rank}] matchBatch {len(a[0])} texts {sum(len(x) for x in a[0])} bytes "
                              f"{(time.perf_counter() - t) * 1e3:.2f} ms  {({k2: round(v2, 2) for k2, v2 in m.timing().items() if v2 > 0.05})} "
                              f"{m.counters()}", file=sys.stderr, flush=True)
                    return r
                return timed
            setattr(m, name, wrap())
        if not sharded:   # rank-local run: hide the process group from dist.py
            saved = dist.is_initialized
            dist.is_initialized = lambda: False
        # size the handle's buffers outside the timed region (a fresh matcher allocates on its first batch): match and
        # pack this rank's share of the first window once; neither call changes the matcher state
        flat = [c for f in files for c in f]
        share = flat[:window][(rank if sharded else 0)::(world if sharded else 1)]
        if share:
            m.matchBatch(share, 32, fetch=False)
            m.packBatch(pd, share, None, loadedPosArr=[m.getLoadingPosition()])
        spent.clear()
        try:
            torch.cuda.synchronize()
            if sharded and world > 1:
                dist.barrier()
            t0 = time.perf_counter()
            streams, stats = compress_speculative(m, files, pd, window, device="cuda" if (sharded and world > 1) else "cpu")
            torch.cuda.synchronize()
            dt = time.perf_counter() - t0
        finally:
            if not sharded:
                dist.is_initialized = saved
        m.close()
        stats["wall_s"] = {k: round(v, 4) for k, v in spent.items()}
        return streams, stats, dt

    run(args.window, True)                     # warm-up (NCCL channels, page cache of the host buffers)
    streams, stats, dt = run(args.window, True)
    dt = max_over_ranks(dt, "cuda" if world > 1 else "cpu")
    line = {"metric": "compression input GB/s (match stage + stream packing, sharded schedule)", "n_gpus": world,
            "value": nbytes / dt / 1e9, "unit": "GB/s", "seconds": dt, "fasta_bytes": nbytes, "genomes": args.genomes,
            "tier": args.tier, "window": args.window, "mode": args.mode, **stats,
            "stream_bytes": {k: len(v) for k, v in streams.items()},
            "stream_md5": {k: hashlib.md5(v).hexdigest() for k, v in streams.items()}}
    if args.check and rank == 0:
        want, wstats, _ = run(1, False)       # one contig per round on one GPU: the sequential order
        line["identical_to_sequential"] = all(streams[k] == want[k] for k in want)
        line["sequential_rounds"] = wstats["rounds"]
    if rank == 0:
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    if args.check and rank == 0 and not line["identical_to_sequential"]:
        raise SystemExit("sharded streams differ from the sequential schedule")


if __name__ == "__main__":
    main()
