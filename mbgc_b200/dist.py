"""Multi-GPU plumbing of the match stage: one process per GPU, torch.distributed for the control plane only.

The path shards by TARGET: within a batch every rank matches its own targets against the same
frozen matcher state (index + reference replicated per GPU), so the data path needs no collective.  What the
ranks do exchange is tiny and host-side: per-target match counts (to lay the lists out in target order, the
order MBGC_Encoder concatenates per-target streams in, mbgccoder/MBGC_Encoder.cpp:542-564) and timings
(max over ranks).  Works with the "nccl" backend on GPUs and with "gloo" on CPU (tests)."""
import numpy as np


def shard_range(n_items, rank, world):
    """contiguous, balanced split of range(n_items): the first n % world ranks get one item more"""
    base, extra = divmod(n_items, world)
    begin = rank * base + min(rank, extra)
    return begin, begin + base + (1 if rank < extra else 0)


def max_over_ranks(value, device="cpu"):
    """max of a python float over all ranks (identity when torch.distributed is not initialised)"""
    import torch
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()):
        return float(value)
    t = torch.tensor([float(value)], dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def sum_over_ranks(value, device="cpu"):
    import torch
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()):
        return int(value)
    t = torch.tensor([int(value)], dtype=torch.int64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.SUM)
    return int(t.item())


def gather_in_target_order(local_lists, n_targets, device="cpu"):
    """local_lists: this rank's per-target (n_i, 3) uint64 match arrays for shard_range(n_targets, rank, world).
    Returns, on every rank, the list of all n_targets arrays in target order."""
    import torch
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()):
        return list(local_lists)
    world, rank = dist.get_world_size(), dist.get_rank()
    counts = torch.zeros(n_targets, dtype=torch.int64, device=device)
    b, e = shard_range(n_targets, rank, world)
    assert e - b == len(local_lists)
    for k, a in enumerate(local_lists):
        counts[b + k] = len(a)
    dist.all_reduce(counts, op=dist.ReduceOp.SUM)
    counts_h = counts.cpu().numpy()
    per_rank = [int(counts_h[slice(*shard_range(n_targets, r, world))].sum()) for r in range(world)]
    width = max(per_rank) if per_rank else 0
    mine = np.zeros((width, 3), dtype=np.int64)
    if local_lists and per_rank[rank]:
        mine[:per_rank[rank]] = np.concatenate(local_lists).astype(np.int64, copy=False)
    bufs = [torch.zeros((width, 3), dtype=torch.int64, device=device) for _ in range(world)]
    dist.all_gather(bufs, torch.from_numpy(mine).to(device))
    out = []
    for r in range(world):
        rb, re_ = shard_range(n_targets, r, world)
        flat = bufs[r].cpu().numpy().astype(np.uint64)
        o = 0
        for t in range(rb, re_):
            out.append(flat[o:o + int(counts_h[t])])
            o += int(counts_h[t])
    return out


# ------------------------------------------------------------------------------------------------------
# The exchange step of the sharded path: reference extensions
#
# Index and reference are replicated.  When a target contig qualifies for reference extension
# (MultipleGenomeMatchingProcessor.cpp:277-287, 389-398, 441) its bytes must reach every replica, and every
# replica must apply the same loadRef calls in the same order: the index build is deterministic in that order
# (insert sequence numbers), so the replicas stay bit-identical without exchanging table contents.
# ------------------------------------------------------------------------------------------------------

def broadcast_bytes(data, src, device="cpu"):
    """Sends a uint8 array from rank `src` to every rank (two collectives: length, payload).  With the nccl backend
    pass device="cuda": the payload travels GPU to GPU over NVLink; gloo uses host tensors."""
    import torch
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()):
        return np.ascontiguousarray(data, dtype=np.uint8)
    rank = dist.get_rank()
    n = torch.tensor([len(data) if rank == src else 0], dtype=torch.int64, device=device)
    dist.broadcast(n, src)
    n = int(n.item())
    if rank == src:
        buf = torch.from_numpy(np.ascontiguousarray(data, dtype=np.uint8)).to(device)
    else:
        buf = torch.empty(n, dtype=torch.uint8, device=device)
    if n:
        dist.broadcast(buf, src)
    return buf.cpu().numpy()


def extend_reference_everywhere(matcher, text, owner, load_rc, device="cpu", separator=0):
    """`owner` holds `text` (others pass None); afterwards every rank's matcher has loaded it (and its reverse
    complement if load_rc), i.e. the replicas made the same loadRef call."""
    data = broadcast_bytes(text, owner, device)
    matcher.loadRef(data, load_rc, True, separator)
    return data


def match_speculative(matcher, texts, extension_rule, window, device="cpu", min_match_length=32):
    """Sequential-equivalent matching of `texts` (in order) with a growing reference, sharded over the ranks by
    optimistic batching.

    The reference's sequential schedule is: match text i, decide from its match list whether it extends the
    reference (and with its reverse complement), load it, go on to i + 1
    (MultipleGenomeMatchingProcessor.cpp:274-287).  A text that does not extend the reference leaves the matcher
    state untouched, so a whole window of texts can be matched in parallel against the frozen state; results are
    committed in order up to and including the first text that extends the reference, that extension is applied on
    every rank (broadcast from the rank that matched it), and the rest of the window is matched again.  On
    near-identical collections nothing qualifies and every window commits whole; on divergent ones this
    degenerates to one text per round.

    Every rank passes the same `texts` (uint8 arrays).  extension_rule(text_len, matches) -> (extend, load_rc) must
    be a pure function of its arguments (the reference's rule needs the unmatched-character count of
    MBGC_Encoder::processMatches, which the caller supplies through this hook).  Returns the per-text match lists,
    identical on every rank and identical to the one-text-at-a-time schedule, and the number of rounds."""
    import torch.distributed as dist
    world = dist.get_world_size() if dist.is_available() and dist.is_initialized() else 1
    rank = dist.get_rank() if world > 1 else 0
    n = len(texts)
    out = [None] * n
    done, rounds = 0, 0
    while done < n:
        hi = min(n, done + window)
        spans = [shard_range(hi - done, r, world) for r in range(world)]
        b, e = spans[rank]
        mine = [texts[done + k] for k in range(b, e)]
        local = matcher.matchBatch(mine, min_match_length) if mine else []
        lists = gather_in_target_order(local, hi - done, device)
        rounds += 1
        committed = 0
        for k, m in enumerate(lists):
            out[done + k] = m
            committed += 1
            extend, load_rc = extension_rule(len(texts[done + k]), m)
            if extend:
                owner = next(r for r, (sb, se) in enumerate(spans) if sb <= k < se)  # the rank that matched it
                extend_reference_everywhere(matcher, texts[done + k] if rank == owner else None, owner, load_rc, device)
                break
        done += committed
    return out, rounds


# ------------------------------------------------------------------------------------------------------
# The whole hot path sharded: match stage + stream packing (K6), reference extensions decided by the reference's
# own rule on the packer's unmatched count.
# ------------------------------------------------------------------------------------------------------

STREAM_NAMES = ("literals", "mapOff", "mapOff5", "mapLen", "gapDelta", "gapFlags")
SEQ_SEPARATOR_MARK = (ord('"') + 128) & 0xFF   # mbgccoder/MBGC_Params.h:46
FILE_SEPARATOR_MARK = (ord(';') + 128) & 0xFF  # mbgccoder/MBGC_Params.h:47


def _gather_packed(local, n_items, device="cpu"):
    """local: this rank's per-text results [(unmatched, [six bytes objects])] for its shard of n_items texts.
    Returns the same for all n_items texts, in text order, on every rank (sizes by all_reduce, payload by all_gather)."""
    import torch
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()):
        return list(local)
    world, rank = dist.get_world_size(), dist.get_rank()
    meta = torch.zeros((n_items, 7), dtype=torch.int64, device=device)  # unmatched + six stream sizes per text
    b, e = shard_range(n_items, rank, world)
    assert e - b == len(local)
    for k, (unmatched, segs) in enumerate(local):
        meta[b + k, 0] = int(unmatched)
        for s_, seg in enumerate(segs):
            meta[b + k, 1 + s_] = len(seg)
    dist.all_reduce(meta, op=dist.ReduceOp.SUM)
    meta_h = meta.cpu().numpy()
    per_rank = [int(meta_h[slice(*shard_range(n_items, r, world)), 1:].sum()) for r in range(world)]
    width = max(per_rank) if per_rank else 0
    mine = np.zeros(width, dtype=np.uint8)
    flat = b"".join(seg for _, segs in local for seg in segs)
    mine[:len(flat)] = np.frombuffer(flat, dtype=np.uint8)
    bufs = [torch.zeros(width, dtype=torch.uint8, device=device) for _ in range(world)]
    dist.all_gather(bufs, torch.from_numpy(mine).to(device))
    out = []
    for r in range(world):
        rb, re_ = shard_range(n_items, r, world)
        data = bufs[r].cpu().numpy().tobytes()
        o = 0
        for t in range(rb, re_):
            segs = []
            for s_ in range(6):
                n = int(meta_h[t, 1 + s_])
                segs.append(data[o:o + n])
                o += n
            out.append((int(meta_h[t, 0]), segs))
    return out


def compress_speculative(matcher, files, pack_params, window, device="cpu", min_match_length=32,
                         unmatched_fraction_factor=128, unmatched_fraction_rc_factor=8, max_pass_bytes=512 << 20):
    """The reference's sequential (-t 1) encoder schedule over `files` (a list of files, each a list of contig texts),
    with both stages of the hot path on the matcher's device and the contigs of a window sharded over the ranks
    (optimistic batching):

        per contig   matchTexts -> processMatches -> if unmatched * factor > length: loadRef(contig [, +RC])
                     (MultipleGenomeMatchingProcessor.cpp:274-287; MGMP_Params.h:179-196)
        per file     FILE_SEPARATOR_MARK in the flag stream, loadSeparator, refExtLoadedPosArr grows
                     (mbgccoder/MBGC_Encoder.cpp:493-509)

    A contig that does not extend the reference leaves index and reference untouched, and the one-byte separators
    appended between files only move the loading cursor over bytes that are zero anyway, so a window of contigs
    (across file boundaries) is matched AND packed in parallel against the frozen state; results are committed in
    order up to and including the first contig that extends the reference; the extension is broadcast and applied on
    every replica and the rest of the window is processed again.  Every rank passes the same `files`.  `window` is the
    largest number of contigs matched speculatively; the actual window follows the commit rate, so a divergent
    collection (every contig extends the reference) costs a few contigs per round instead of the whole window.

    Returns ({stream name: bytes}, stats) -- identical on every rank and byte-identical to the sequential schedule
    (tests/seq_schedule.py::encode_streams, which reproduces the reference's own stream dumps).  The literal stream
    does not include the G0 bootstrap the encoder writes before the first target (MBGC_Encoder.cpp:34-37)."""
    import torch.distributed as dist
    world = dist.get_world_size() if dist.is_available() and dist.is_initialized() else 1
    rank = dist.get_rank() if world > 1 else 0
    flat = [(fi, ci) for fi, f in enumerate(files) for ci in range(len(f))]
    last_of_file = {(fi, len(f) - 1) for fi, f in enumerate(files) if len(f)}
    out = {n: [] for n in STREAM_NAMES}
    loaded = [matcher.getLoadingPosition()]
    file_start = matcher.getLoadedRefLength()
    stats = dict(rounds=0, extensions=0, unmatched=0, contigs=len(flat), repacked=0)
    # files without contigs still close with a separator
    def close_file():
        nonlocal file_start
        if pack_params.enableExtensionsWithMismatches:
            out["gapFlags"].append(bytes([FILE_SEPARATOR_MARK]))
        matcher.loadSeparator(0)
        loaded.append(loaded[-1] + (matcher.getLoadedRefLength() - file_start))
        file_start = matcher.getLoadedRefLength()

    next_file = 0
    done = 0
    cur_window = window
    while done < len(flat):
        hi = min(len(flat), done + cur_window)
        spans = [shard_range(hi - done, r, world) for r in range(world)]
        b, e = spans[rank]
        mine = [files[fi][ci] for fi, ci in flat[done + b:done + e]]
        local = []
        at = 0
        while at < len(mine):   # sub-batches that fit one device pass (the lists stay on the device between the two calls)
            end, nbytes = at, 0
            while end < len(mine) and (end == at or nbytes + len(mine[end]) <= max_pass_bytes):
                nbytes += len(mine[end])
                end += 1
            part = mine[at:end]
            matcher.matchBatch(part, min_match_length, fetch=False)
            res = matcher.packBatch(pack_params, part, None, loadedPosArr=loaded)
            local += [(int(res["unmatched"][k]), [res[n][k] for n in STREAM_NAMES]) for k in range(len(part))]
            at = end
        packed = _gather_packed(local, hi - done, device)
        stats["rounds"] += 1
        committed = 0
        for k, (unmatched, segs) in enumerate(packed):
            fi, ci = flat[done + k]
            while next_file < fi:          # (empty files in between)
                close_file()
                next_file += 1
            text = files[fi][ci]
            for n, seg in zip(STREAM_NAMES, segs):
                out[n].append(seg)
            out["literals"].append(bytes([SEQ_SEPARATOR_MARK]))
            stats["unmatched"] += unmatched
            committed += 1
            extend = unmatched * unmatched_fraction_factor > len(text)
            if extend:
                owner = next(r for r, (sb, se) in enumerate(spans) if sb <= k < se)
                extend_reference_everywhere(matcher, text if rank == owner else None, owner,
                                            unmatched * unmatched_fraction_rc_factor > len(text), device)
                stats["extensions"] += 1
            if (fi, ci) in last_of_file:
                close_file()
                next_file = fi + 1
            if extend:
                break
        stats["repacked"] += (hi - done) - committed
        done += committed
        # speculation that fails is wasted work: follow the commit rate (twice what the last round committed, at least
        # one contig per rank, at most the caller's window); identical on every rank, so the schedule stays in lock step
        cur_window = max(world, min(window, 2 * committed))
    while next_file < len(files):
        close_file()
        next_file += 1
    return {n: b"".join(v) for n, v in out.items()}, stats


# ------------------------------------------------------------------------------------------------------
# The collection pass (mbgcg_seq_*, C++: mbgc_b200/csrc/collection.cpp) sharded over ranks: every rank reads ONLY its
# own files and keeps its stream bytes; what travels is one integer per rank.
# ------------------------------------------------------------------------------------------------------

class SpeculationFailed(RuntimeError):
    """a rank other than the last extended the reference: the shards behind it were matched against a stale state"""


def compress_sharded_files(matcher, paths, cfg, device="cpu", readers=None, keep_streams=True):
    """The sequential (-t 1) schedule of `paths` (files in target order; every rank passes the same list) with the files
    split into one contiguous shard per rank.  Rank r runs the collection pass on its shard against the state the
    sequential schedule has after the files in front of it PROVIDED none of them extends the reference: one region
    separator per file (mbgccoder/MBGC_Encoder.cpp:498-505) and nothing else.  That speculation is checked afterwards
    with one all_gather of the per-rank extension counts; it holds for near-identical collections (the README's
    1k-E.-coli regime) and raises SpeculationFailed otherwise (use compress_speculative, or one process
    with mbgcg_create_multi, for divergent collections).

    Returns (streams of this rank's shard, info, extensions per rank).  The sequential schedule's streams are the
    per-rank streams concatenated in rank order; nothing but the counts is exchanged."""
    import torch
    import torch.distributed as dist
    world = dist.get_world_size() if dist.is_available() and dist.is_initialized() else 1
    rank = dist.get_rank() if world > 1 else 0
    a, b = shard_range(len(paths), rank, world)
    loaded = [matcher.getLoadingPosition()]
    if cfg.separatorPerFile:
        start = matcher.getLoadedRefLength()
        for _ in range(a):
            matcher.loadSeparator(0)
            now = matcher.getLoadedRefLength()
            loaded.append(loaded[-1] + (now - start))
            start = now
    streams, info = matcher.compressFiles(paths[a:b], cfg, loadedPosArr=loaded, readers=readers, keep_streams=keep_streams)
    ext = [info["extensions"]]
    if world > 1:
        t = torch.zeros(world, dtype=torch.int64, device=device)
        t[rank] = info["extensions"]
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
        ext = [int(x) for x in t.cpu().tolist()]
    if any(ext[:-1]):
        raise SpeculationFailed(f"reference extensions per rank {ext}: shards behind the first extending rank are stale")
    return streams, info, ext
