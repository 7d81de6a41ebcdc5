"""Emulation, on the oracle entry points, of the batch-frozen schedule the collection pass runs for MBGC's PARALLEL mode
(mbgc_b200/csrc/collection.cpp, SeqConfig::parallelSchedule).  Test infrastructure.

The reference's parallel mode has no deterministic schedule of its own; what is fixed is the per-target
work (paths relative to /root/reference):
  MultipleGenomeMatchingProcessor.cpp:340-429   processTarget: lock, per contig matchTexts(lock) -> processMatches(lock) ->
                                                targetRefExtensions[i] += contig (+ its reverse complement)
  MultipleGenomeMatchingProcessor.cpp:433-468   finalize in target order: loadRef(extension), hook, release lock
  mbgccoder/MBGC_Encoder.cpp:202-205, 542-564   dissimilar-contig rule; per-target stream concatenation, separator
Here: the targets of a window are matched against the state at the window's start; the window grouping is an input
(taken from the device pass), everything else is recomputed with the CPU oracle."""
import ctypes as C

import numpy as np

from seq_schedule import STREAMS, reference_factor  # noqa: F401
from util import OracleMatcher, PackParams, oracle_lib, read_fasta, u8p

SEQ_SEPARATOR_MARK = (ord('"') + 128) & 0xFF
FILE_SEPARATOR_MARK = (ord(';') + 128) & 0xFF


def encode_streams_parallel(paths, mode, limit, windows, dis_outrun, tweak, min_len_dissimilar=1024):
    """windows: list of lists of file indices (in order).  Initial reference = first contig of file 0 + its reverse
    complement (as in seq_schedule).  Returns ({stream: bytes}, stats)."""
    lib = oracle_lib()
    recs = [read_fasta(p) for p in paths]
    skip_margin = 24 if mode >= 2 else 16
    rc_factor = 128 if mode >= 2 else 8
    m = OracleMatcher(limit, 32, 16, 1, skip_margin)
    m.disable_sliding_window()
    pp = PackParams.for_mode(mode)
    m.load_ref(recs[0][0][1], rc=True, add_sep=True)
    loaded = [m.loading_position]
    out = {n: [] for n in STREAMS}
    stats = dict(matches=0, unmatched=0, reruns=0, locks=[], ref_ext=[])
    finalized = 0

    def run(contigs, locks):
        """match + pack every (file, contig) of `contigs` against the current state -> per contig (unmatched, n, segments)"""
        res = {}
        arr = (C.c_uint64 * len(loaded))(*loaded)
        for (f, k) in contigs:
            a = np.frombuffer(recs[f][k][1], dtype=np.uint8)
            mt = np.ascontiguousarray(m.match(a, 32, lock=locks[f]))
            n = C.c_uint64(len(mt))
            st = lib.mbo_streams_create()
            unmatched = lib.mbo_process_matches(st, C.byref(pp), m.h, mt.ctypes.data_as(C.c_void_p), C.byref(n),
                                                a.ctypes.data_as(u8p), a.size, locks[f], arr, len(loaded))
            segs = []
            for i in range(6):
                sz = lib.mbo_stream_size(st, i)
                segs.append(bytes(np.ctypeslib.as_array(C.cast(lib.mbo_stream_data(st, i), u8p), shape=(sz,))) if sz else b"")
            lib.mbo_streams_destroy(st)
            res[(f, k)] = (unmatched, len(mt), segs)
        return res

    for win in windows:
        locks = {f: m.acquire_lock() for f in win}
        order = [(f, k) for f in win for k in range(len(recs[f]))]
        res = run(order, locks)
        first_target = finalized
        for f in win:
            ext = []
            for k in range(len(recs[f])):
                seq = recs[f][k][1]
                unmatched, nm, segs = res[(f, k)]
                if first_target + dis_outrun < f and len(seq) > min_len_dissimilar and unmatched * (128 // tweak) > len(seq):
                    rest = order[order.index((f, k)):]
                    res.update(run(rest, locks))
                    stats["reruns"] += len(rest)
                    first_target = finalized
                    unmatched, nm, segs = res[(f, k)]
                for name, seg in zip(STREAMS, segs):
                    out[name].append(seg)
                out["literals"].append(bytes([SEQ_SEPARATOR_MARK]))
                stats["matches"] += nm
                stats["unmatched"] += unmatched
                if unmatched * 128 > len(seq):
                    ext.append(seq)
                    if unmatched * rc_factor > len(seq):
                        rc = np.empty(len(seq), dtype=np.uint8)
                        a = np.frombuffer(seq, dtype=np.uint8)
                        lib.mbo_upper_reverse_complement(a.ctypes.data_as(u8p), len(seq), rc.ctypes.data_as(u8p))
                        ext.append(rc.tobytes())
            out["gapFlags"].append(bytes([FILE_SEPARATOR_MARK]))
            start = m.loaded_ref_length
            if ext:
                m.load_ref(b"".join(ext), rc=False, add_sep=True)
            m.load_separator(0)
            stats["ref_ext"].append(m.loaded_ref_length - start)
            loaded.append(m.loaded_ref_length)
            stats["locks"].append(locks[f])
            m.release_lock(locks[f])
            finalized += 1
    stats["ref_length"] = m.ref_length
    m.close()
    return {n: b"".join(v) for n, v in out.items()}, stats
