"""Emulation of MBGC's deterministic (-t 1) encoder schedule on top of the oracle entry points, producing the
six match/literal streams the reference hands to its back-end coders.  Test infrastructure.

Follows (paths relative to /root/reference):
  MultipleGenomeMatchingProcessor.cpp:66-150   loadG0Ref, sequential branch: only the first contig of file 0
                                               seeds the reference; file 0 is then a target like the others
  MultipleGenomeMatchingProcessor.cpp:152-192  initMatcher: limit = referenceFactor * max(fileSize0, 2 MiB) * 2
  MultipleGenomeMatchingProcessor.cpp:232-313  processTargetsWithParallelIO (per contig: matchTexts,
                                               processMatches, extension decision, loadRef; per file: separator)
  mbgccoder/MBGC_Encoder.cpp:34-37,489-509,786-791   literal stream bootstrap, separators, refExtLoadedPosArr
"""
import ctypes as C
import os

import numpy as np

from util import NO_LOCK, OracleMatcher, PackParams, oracle_lib, read_fasta, u64p, u8p

STREAMS = ("literals", "mapOff", "mapOff5", "mapLen", "gapDelta", "gapFlags")
SEQ_SEPARATOR_MARK = (ord('"') + 128) & 0xFF


def reference_factor(n_files):
    clz = 32 - int(n_files).bit_length()
    tmp = 15 - clz // 3
    return 1 << (5 if tmp < 5 else (12 if tmp > 12 else tmp))


class LitRange(C.Structure):
    _fields_ = [("start", C.c_uint64), ("len", C.c_uint64)]


def encode_streams(paths, mode, on_match=None, lit_ext=None):
    """Returns ({stream name: bytes}, stats).  on_match(contig_bytes, matcher, oracle_matches) lets a caller
    cross-check another matcher (e.g. the CUDA one) in lock step.  lit_ext = (refLiteralMinimalLengthExt,
    refLiteralBeforeAfterExt) switches MBGC_Encoder::processLiteral on (the reference's developer option -N / -F): a contig
    that is not loaded as a whole extends the reference with its long literals (MultipleGenomeMatchingProcessor.cpp:281-286)."""
    lib = oracle_lib()
    recs = [read_fasta(p) for p in paths]
    c0 = recs[0][0][1]
    basic = max(os.path.getsize(paths[0]), 1 << 21)
    limit = reference_factor(len(paths)) * basic * 2
    limit = min(limit, 2**32 - 1)
    limit = max(limit, len(c0))
    skip_margin = 24 if mode >= 2 else 16
    rc_factor = 128 if mode >= 2 else 8
    m = OracleMatcher(limit, 32, 16, 1, skip_margin)
    m.disable_sliding_window()
    pp = PackParams.for_mode(mode)
    streams = lib.mbo_streams_create()
    lit0 = np.frombuffer(c0 + bytes([SEQ_SEPARATOR_MARK]), dtype=np.uint8)
    m.load_ref(c0, rc=True, add_sep=True)
    loaded = [m.loading_position]
    stats = dict(matches=0, unmatched=0, limit=limit)
    events = []
    for recs_i in recs:
        start = m.loaded_ref_length
        for _, seq in recs_i:
            a = np.frombuffer(seq, dtype=np.uint8)
            mt = m.match(a, 32)
            if on_match is not None:
                on_match(a, m, mt)
            stats["matches"] += len(mt)
            mt = np.ascontiguousarray(mt)
            n = C.c_uint64(len(mt))
            arr = (C.c_uint64 * len(loaded))(*loaded)
            if lit_ext is None:
                unmatched = lib.mbo_process_matches(streams, C.byref(pp), m.h, mt.ctypes.data_as(C.c_void_p), C.byref(n),
                                                    a.ctypes.data_as(u8p), a.size, NO_LOCK, arr, len(loaded))
                ranges = []
            else:
                cap = len(mt) + 2
                rbuf = (LitRange * cap)()
                nr = C.c_uint64(0)
                lib.mbo_process_matches_lit.restype = C.c_uint64
                unmatched = lib.mbo_process_matches_lit(streams, C.byref(pp), m.h, mt.ctypes.data_as(C.c_void_p), C.byref(n),
                                                        a.ctypes.data_as(u8p), C.c_uint64(a.size), C.c_uint64(NO_LOCK), arr,
                                                        C.c_uint64(len(loaded)), C.c_uint64(lit_ext[0]), C.c_uint64(lit_ext[1]),
                                                        rbuf, C.c_uint64(cap), C.byref(nr))
                ranges = [(rbuf[k].start, rbuf[k].len) for k in range(nr.value)]
                stats.setdefault("lit_ranges", []).append(ranges)
            stats["unmatched"] += unmatched
            if unmatched * 128 > a.size:
                m.load_ref(a, rc=unmatched * rc_factor > a.size, add_sep=True)
            elif ranges:
                ext = np.concatenate([a[s0:s0 + ln] for s0, ln in ranges])
                m.load_ref(ext, rc=unmatched * rc_factor > a.size, add_sep=True)
            lib.mbo_after_sequence(streams)
        lib.mbo_after_target(streams, C.byref(pp))
        m.load_separator(0)
        loaded.append(loaded[-1] + (m.loaded_ref_length - start))
    out = {}
    for i, name in enumerate(STREAMS):
        n = lib.mbo_stream_size(streams, i)
        p = lib.mbo_stream_data(streams, i)
        out[name] = bytes(np.ctypeslib.as_array(C.cast(p, u8p), shape=(n,))) if n else b""
    out["literals"] = bytes(lit0) + out["literals"]
    lib.mbo_streams_destroy(streams)
    stats["ref_length"] = m.ref_length
    m.close()
    return out, stats
