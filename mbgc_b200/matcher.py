"""ctypes binding of libmbgc_b200.so (include/mbgc_b200.h) with the reference's method names.

Reference interface mirrored: /root/reference/matching/SlidingWindowSparseEMMatcher.h:86-126
(loadRef, loadSeparator, matchTexts, acquire/releaseWorkerMatchingLockPos, getRef, getRefLength, ...).
"""
import ctypes as C
import os
import time

import numpy as np

NO_LOCK = 2**64 - 1  # SW_END_ERASED_FLAG

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None

u8p = C.POINTER(C.c_uint8)
u64p = C.POINTER(C.c_uint64)

T_NAMES = ("h2d", "probe", "compact", "extend", "resolve", "d2h", "total", "anchor", "clean", "fullprobe", "memo")


class MatcherError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__(f"mbgc_b200 error {code}: {msg}")
        self.code = code


class _State(C.Structure):
    _fields_ = [("pos1", C.c_int64), ("maxRefLength", C.c_uint64), ("reachedRefLengthCount", C.c_int32),
                ("circularBuffer", C.c_int32), ("swSize", C.c_uint64), ("swEnd", C.c_uint64),
                ("samplingPos", C.c_uint64), ("hashSize", C.c_uint32), ("K", C.c_int32)]


STREAM_NAMES = ("literals", "mapOff", "mapOff5", "mapLen", "gapDelta", "gapFlags")  # MBGCG_S_*


class PackParams(C.Structure):
    """mbgcg_pack_params_t: the MBGC_Params members MBGC_Encoder::processMatches reads
    (/root/reference/mbgccoder/MBGC_Params.h:66-97; presets :886-922)."""
    _fields_ = [(n, C.c_int32) for n in (
        "enableExtensionsWithMismatches", "gapDepthOffsetEncoding", "gapDepthMismatchesEncoding",
        "lazyDecompressionSupport", "frugal64bitLenEncoding", "enable40bitReference", "mismatchesWithExclusion",
        "mmsMismatchesInitialScore", "mmsMismatchesScoreThreshold", "mmsMismatchPenalty", "mmsMatchBonus",
        "reserved")] + [("gapBreakingMatchMinLength", C.c_uint64)]

    @staticmethod
    def for_mode(mode, enable40bit=False, lazy=True):
        """defaults of the reference for `mbgc c -m mode` (mode 0 writes plain u32 lengths, MBGC_Params.h:900)"""
        return PackParams(1, 64, 2, 1 if lazy else 0, 0 if mode == 0 else 1, 1 if enable40bit else 0, 1,
                          125, 500, 50, 50, 0, 256)


class DecodeInput(C.Structure):
    """mbgcg_decode_input_t"""
    _fields_ = [("stream", C.c_void_p * 6), ("streamLen", C.c_uint64 * 6), ("refTotalLength", C.c_uint64), ("g0IsTarget", C.c_int32),
                ("g0SeqCount", C.c_uint32), ("rcInReference", C.c_int32), ("contigsIndividuallyReversed", C.c_int32),
                ("nTargets", C.c_uint32), ("seqCounts", C.POINTER(C.c_uint32)), ("unmatchedFractionFactors", C.POINTER(C.c_uint8)),
                ("lockPos", C.POINTER(C.c_uint64))]


class DecodeResult(C.Structure):
    """mbgcg_decode_result_t"""
    _fields_ = [("nContigs", C.c_uint64), ("contigLen", C.POINTER(C.c_uint64)), ("contigUnmatched", C.POINTER(C.c_uint64)),
                ("dnaLen", C.c_uint64), ("refPos", C.c_uint64), ("reachedRefLengthCount", C.c_int32), ("batches", C.c_uint32),
                ("parseMs", C.c_float), ("copyMs", C.c_float), ("launches", C.c_uint64)]


class SeqConfig(C.Structure):
    """mbgcg_seq_config_t: what the sequential target schedule reads from MGMP_Params / MBGC_Params"""
    _fields_ = [("pack", PackParams), ("minMatchLength", C.c_uint32), ("unmatchedFractionFactor", C.c_int32),
                ("unmatchedFractionRCFactor", C.c_int32), ("rcInReferenceDisabled", C.c_int32),
                ("refRegionSeparators", C.c_int32), ("separatorPerFile", C.c_int32), ("regionSeparator", C.c_char),
                ("chunkBytes", C.c_uint64), ("maxWindowContigs", C.c_uint32), ("parallelSchedule", C.c_int32),
                ("contigsIndividuallyReversed", C.c_int32), ("allowedTargetsOutrun", C.c_int32),
                ("allowedTargetsOutrunForDissimilarContigs", C.c_int32),
                ("unmatchedFractionFactorTweakForDissimilarContigs", C.c_int32),
                ("minimalLengthForDissimilarContigs", C.c_uint64)]

    @staticmethod
    def for_mode(mode, chunk_bytes=0, max_window=0, lazy=True, parallel=False, outrun=64):
        """`mbgc c -m mode` defaults: extension when unmatched * 128 > length, its reverse complement too when
        unmatched * 8 > length (modes 2/3: * 128), MGMP_Params.h:41-42, MBGC_Params.h:906-909.  parallel=True: the
        batch-frozen schedule of the reference's parallel mode with the mode's dissimilar-contig settings
        (MGMP_Params.h:57-63, MBGC_Params.h:893-905)."""
        dis_outrun, tweak = {0: (4, 32), 2: (0, 2)}.get(mode, (1, 16))
        return SeqConfig(PackParams.for_mode(mode, lazy=lazy), 32, 128, 128 if mode >= 2 else 8, 0, 1, 1 if lazy else 0,
                         b"\0", chunk_bytes, max_window, 1 if parallel else 0, 1, outrun, dis_outrun, tweak, 1024)


class SeqEvent(C.Structure):
    _fields_ = [("type", C.c_int32), ("status", C.c_int32), ("file", C.c_uint64), ("hdr", C.c_void_p),
                ("hdrLen", C.c_uint64), ("seqLen", C.c_uint64), ("nMatches", C.c_uint64), ("kept", C.c_uint64),
                ("unmatched", C.c_uint64), ("extended", C.c_int32), ("extendedRC", C.c_int32), ("contigLock", C.c_uint64), ("window", C.c_uint64),
                ("seg", C.c_void_p * 6), ("segLen", C.c_uint64 * 6), ("nContigs", C.c_uint64),
                ("dnaLineLen", C.c_uint64), ("refExtSize", C.c_uint64), ("lockPos", C.c_uint64),
                ("loadedRefLength", C.c_uint64)]


class SeqStats(C.Structure):
    _fields_ = [(n, C.c_uint64) for n in ("files", "contigs", "windows", "runs", "rerunContigs", "extensions",
                                          "imageBytes", "textBytes", "pageableBytes", "launches", "chunks")] + \
               [(n, C.c_double) for n in ("ingestMs", "matchMs", "packMs", "loadRefMs", "waitStageMs")] + [("wireBytes", C.c_uint64)]


class SeqResult(C.Structure):
    _fields_ = [("stream", C.c_void_p * 6), ("streamLen", C.c_uint64 * 6), ("nFiles", C.c_uint64), ("nContigs", C.c_uint64),
                ("seqCounts", C.c_void_p), ("dnaLineLens", C.c_void_p), ("refExtSizes", C.c_void_p), ("lockPos", C.c_void_p),
                ("status", C.c_void_p), ("headers", C.c_void_p), ("headersLen", C.c_uint64), ("matches", C.c_uint64),
                ("unmatched", C.c_uint64), ("extensions", C.c_uint64), ("digest", C.c_uint64)]


NEXT_IMAGE_FN = C.CFUNCTYPE(C.c_int, C.c_void_p, C.c_int, C.POINTER(C.c_void_p), C.POINTER(C.c_uint64))
EV_CONTIG, EV_FILE_END, EV_NEED_INPUT, EV_DONE = 1, 2, 3, 4
SEQ_SEPARATOR_MARK = (ord('"') + 128) & 0xFF   # mbgccoder/MBGC_Params.h:46
FILE_SEPARATOR_MARK = (ord(';') + 128) & 0xFF  # mbgccoder/MBGC_Params.h:47


class _PackResult(C.Structure):
    _fields_ = [("stream", C.c_void_p * 6), ("streamOff", C.c_void_p * 6), ("unmatched", C.c_void_p),
                ("kept", C.c_void_p)]


def lib_path():
    return os.path.join(_HERE, "libmbgc_b200.so")


def load_library():
    """Loads libmbgc_b200.so.  Raises if it has not been built: the matcher has no other implementation."""
    global _LIB
    if _LIB is not None:
        return _LIB
    path = lib_path()
    if not os.path.exists(path):
        raise MatcherError(-1, f"{path} is missing -- build it with `python -m mbgc_b200.build` "
                               "(there is no CPU fallback for the matcher)")
    L = C.CDLL(path)
    vp = C.c_void_p
    L.mbgcg_create.restype = C.c_int
    L.mbgcg_create.argtypes = [C.POINTER(vp), C.c_uint64, C.c_uint32, C.c_int, C.c_int, C.c_int, C.c_int]
    L.mbgcg_create_multi.restype = C.c_int
    L.mbgcg_create_multi.argtypes = [C.POINTER(vp), C.c_uint64, C.c_uint32, C.c_int, C.c_int, C.c_int, C.POINTER(C.c_int), C.c_int]
    L.mbgcg_device_count.restype = C.c_int
    L.mbgcg_device_count.argtypes = [vp]
    L.mbgcg_seq_begin.restype = C.c_int
    L.mbgcg_seq_begin.argtypes = [vp, C.POINTER(SeqConfig), u64p, C.c_uint64]
    L.mbgcg_seq_push_file.restype = C.c_int
    L.mbgcg_seq_push_file.argtypes = [vp, vp, C.c_uint64]
    L.mbgcg_seq_finish_input.restype = C.c_int
    L.mbgcg_seq_finish_input.argtypes = [vp]
    L.mbgcg_seq_wants_input.restype = C.c_int
    L.mbgcg_seq_wants_input.argtypes = [vp]
    L.mbgcg_seq_next.restype = C.c_int
    L.mbgcg_seq_next.argtypes = [vp, C.POINTER(SeqEvent)]
    L.mbgcg_seq_run.restype = C.c_int
    L.mbgcg_seq_run.argtypes = [vp, NEXT_IMAGE_FN, vp, C.c_int, C.POINTER(SeqResult)]
    L.mbgcg_seq_end.restype = C.c_int
    L.mbgcg_seq_end.argtypes = [vp, C.POINTER(SeqStats)]
    L.mbgcg_image_alloc.restype = vp
    L.mbgcg_image_alloc.argtypes = [vp, C.c_uint64]
    L.mbgcg_image_max.restype = C.c_uint64
    L.mbgcg_image_max.argtypes = [vp]
    L.mbgcg_image_pack_bound.restype = C.c_uint64
    L.mbgcg_image_pack_bound.argtypes = [C.c_uint64]
    L.mbgcg_image_pack.restype = C.c_int
    L.mbgcg_image_pack.argtypes = [vp, C.c_uint64, vp, C.c_uint64, u64p]
    L.mbgcg_image_pack_file.restype = C.c_int
    L.mbgcg_image_pack_file.argtypes = [C.c_char_p, vp, C.c_uint64, u64p, u64p]
    L.mbgcg_seq_run_files.restype = C.c_int
    L.mbgcg_seq_run_files.argtypes = [vp, C.POINTER(C.c_char_p), C.c_uint64, C.c_int, C.c_int, C.c_int, C.POINTER(SeqResult),
                                      C.POINTER(C.c_double)]
    L.mbgcg_seq_push_packed.restype = C.c_int
    L.mbgcg_seq_push_packed.argtypes = [vp, vp]
    L.mbgcg_image_free.restype = C.c_int
    L.mbgcg_image_free.argtypes = [vp, vp]
    L.mbgcg_set_host_mirror.argtypes = [vp, C.c_int]
    L.mbgcg_rc_match_sequence.restype = C.c_int
    L.mbgcg_rc_match_sequence.argtypes = [vp, vp, C.c_uint64, C.c_uint32, C.c_uint32, u64p, vp, u64p, vp, u64p, C.POINTER(C.c_float)]
    L.mbgcg_decode_targets.restype = C.c_int
    L.mbgcg_decode_targets.argtypes = [vp, C.POINTER(PackParams), C.POINTER(DecodeInput), C.POINTER(DecodeResult)]
    L.mbgcg_decode_fetch.restype = C.c_int
    L.mbgcg_decode_fetch.argtypes = [vp, C.c_uint64, C.c_uint64, vp]
    L.mbgcg_decode_device_dna.restype = vp
    L.mbgcg_decode_device_dna.argtypes = [vp, C.c_uint64]
    L.mbgcg_decode_fetch_ref.restype = C.c_int
    L.mbgcg_decode_fetch_ref.argtypes = [vp, C.c_uint64, C.c_uint64, vp]
    L.mbgcg_symbols_pack.restype = C.c_int
    L.mbgcg_symbols_pack.argtypes = [vp, C.c_char_p, C.c_int, vp, C.c_uint64, vp, u64p]
    L.mbgcg_symbols_unpack.restype = C.c_int
    L.mbgcg_symbols_unpack.argtypes = [vp, C.c_char_p, C.c_int, vp, C.c_uint64, C.c_uint64, vp]
    L.mbgcg_varlen_encode.restype = C.c_int
    L.mbgcg_varlen_encode.argtypes = [vp, C.c_char_p, vp, C.c_uint64, vp, C.c_uint64, u64p]
    L.mbgcg_varlen_decode.restype = C.c_int
    L.mbgcg_varlen_decode.argtypes = [vp, C.c_char_p, vp, C.c_uint64, vp, C.c_uint64]
    L.mbgcg_destroy.argtypes = [vp]
    L.mbgcg_last_error.restype = C.c_char_p
    L.mbgcg_last_error.argtypes = [vp]
    L.mbgcg_create_error.restype = C.c_char_p
    L.mbgcg_disable_sliding_window.argtypes = [vp]
    L.mbgcg_disable_circular_buffer.argtypes = [vp]
    L.mbgcg_set_sliding_window_size.argtypes = [vp, C.c_int]
    L.mbgcg_load_ref.restype = C.c_int
    L.mbgcg_load_ref.argtypes = [vp, vp, C.c_uint64, C.c_int, C.c_int, C.c_char]
    L.mbgcg_load_separator.restype = C.c_int
    L.mbgcg_load_separator.argtypes = [vp, C.c_char]
    L.mbgcg_match.restype = C.c_int
    L.mbgcg_match.argtypes = [vp, vp, C.c_uint64, C.c_uint32, C.c_uint64, C.POINTER(vp), u64p]
    L.mbgcg_match_batch.restype = C.c_int
    L.mbgcg_match_batch.argtypes = [vp, C.c_int, C.POINTER(vp), u64p, C.c_uint32, u64p, C.POINTER(vp), C.POINTER(vp)]
    L.mbgcg_match_batch_device.restype = C.c_int
    L.mbgcg_match_batch_device.argtypes = [vp, vp, C.c_int, u64p, u64p, C.c_uint32, u64p, C.POINTER(vp), C.POINTER(vp)]
    L.mbgcg_pack_batch.restype = C.c_int
    L.mbgcg_pack_batch.argtypes = [vp, C.POINTER(PackParams), C.c_int, C.POINTER(vp), u64p, vp, u64p, u64p, u64p,
                                   C.c_uint64, C.POINTER(_PackResult)]
    L.mbgcg_pack_kept_matches.restype = C.c_int
    L.mbgcg_pack_kept_matches.argtypes = [vp, C.POINTER(vp), u64p]
    L.mbgcg_set_literal_ref_ext.restype = C.c_int
    L.mbgcg_set_literal_ref_ext.argtypes = [vp, C.c_uint64, C.c_uint64, C.c_int]
    L.mbgcg_pack_literal_ref_ext.restype = C.c_int
    L.mbgcg_pack_literal_ref_ext.argtypes = [vp, C.POINTER(vp), u64p]
    L.mbgcg_pack_timing.argtypes = [vp, C.POINTER(C.c_float), u64p, C.POINTER(C.c_int)]
    L.mbgcg_ingest_fasta.restype = C.c_int
    L.mbgcg_ingest_fasta.argtypes = [vp, vp, C.c_uint64, C.POINTER(vp), u64p, C.POINTER(vp), u64p, C.POINTER(C.c_int)]
    L.mbgcg_ingest_timing.argtypes = [vp, C.POINTER(C.c_float), u64p]
    L.mbgcg_copy_ingested.restype = C.c_int
    L.mbgcg_copy_ingested.argtypes = [vp, C.c_uint64, vp]
    L.mbgcg_acquire_lock.restype = C.c_uint64
    L.mbgcg_acquire_lock.argtypes = [vp]
    L.mbgcg_release_lock.restype = C.c_int
    L.mbgcg_release_lock.argtypes = [vp, C.c_uint64]
    L.mbgcg_host_ref.restype = vp
    L.mbgcg_host_ref.argtypes = [vp]
    for f in ("mbgcg_max_ref_length", "mbgcg_ref_length", "mbgcg_loading_position", "mbgcg_loaded_ref_length",
              "mbgcg_total_launches"):
        getattr(L, f).restype = C.c_uint64
        getattr(L, f).argtypes = [vp]
    L.mbgcg_set_position.argtypes = [vp, C.c_uint64, C.c_int]
    L.mbgcg_get_state.argtypes = [vp, C.POINTER(_State)]
    L.mbgcg_set_state.argtypes = [vp, C.POINTER(_State)]
    L.mbgcg_copy_table.restype = C.c_int
    L.mbgcg_copy_table.argtypes = [vp, vp, C.c_uint64]
    L.mbgcg_copy_device_ref.restype = C.c_int
    L.mbgcg_copy_device_ref.argtypes = [vp, C.c_uint64, C.c_uint64, vp]
    L.mbgcg_copy_staged_text.restype = C.c_int
    L.mbgcg_copy_staged_text.argtypes = [vp, C.c_uint64, vp]
    L.mbgcg_selftest_pack.restype = C.c_int
    L.mbgcg_selftest_pack.argtypes = [vp, C.c_uint64, vp]
    L.mbgcg_last_timing.argtypes = [vp, C.POINTER(C.c_float)]
    L.mbgcg_set_stage_timing.argtypes = [vp, C.c_int]
    L.mbgcg_last_counters.argtypes = [vp, u64p]
    L.mbgcg_last_wire_bytes.argtypes = [vp, u64p]
    L.mbgcg_load_ref_timing.argtypes = [vp, C.POINTER(C.c_float), u64p]
    L.mbgcg_stream.restype = vp
    L.mbgcg_stream.argtypes = [vp]
    _LIB = L
    return L


def _u64_array(v):
    """ctypes view of a sequence of uint64 for the C ABI; a C-contiguous uint64 numpy array is passed as is"""
    if isinstance(v, np.ndarray) and v.dtype == np.uint64 and v.flags.c_contiguous:
        return v.ctypes.data_as(u64p)
    return (C.c_uint64 * len(v))(*v)


def _as_u8(buf):
    if isinstance(buf, np.ndarray):
        return np.ascontiguousarray(buf, dtype=np.uint8)
    return np.frombuffer(bytes(buf), dtype=np.uint8)


class Matcher:
    """SlidingWindowSparseEMMatcher on a B200.  Match lists come back as uint64 arrays of shape (n, 3) with
    columns (posSrcText, length, posDestText) -- the TextMatch fields of matching/TextMatchers.h:9-17."""

    def __init__(self, ref_length_limit, L=32, k1=16, k2=1, skip_margin=16, device=0, devices=None):
        """devices: a list of CUDA device indices -> the same matcher replicated on each (mbgcg_create_multi); the
        collection pass (compressSequential) shards its windows over them, everything else runs on the first."""
        self.lib = load_library()
        h = C.c_void_p()
        if devices is None:
            rc = self.lib.mbgcg_create(C.byref(h), ref_length_limit, L, k1, k2, skip_margin, device)
        else:
            devs = (C.c_int * len(devices))(*devices)
            rc = self.lib.mbgcg_create_multi(C.byref(h), ref_length_limit, L, k1, k2, skip_margin, devs, len(devices))
        if rc != 0:
            raise MatcherError(rc, self.lib.mbgcg_create_error().decode())
        self.h = h

    # -- row f3: SimpleSequenceMatcher::rcMatchSequence (matching/SimpleSequenceMatcher.cpp:165-176) ------------------------------
    def rcMatchSequence(self, sequence, targetMatchLength=55, minMatchLength=2**32 - 1):
        """-> (rewritten sequence, rcMapOff, rcMapLen, {"index_ms", "query_ms"}) as uint8 arrays"""
        s = _as_u8(sequence).copy()
        cap = s.size // 24 + 4
        off = np.zeros(4 * cap, dtype=np.uint8)
        ln = np.zeros(10 * cap, dtype=np.uint8)
        n, no, nl = C.c_uint64(), C.c_uint64(), C.c_uint64()
        ms = (C.c_float * 2)()
        self._check(self.lib.mbgcg_rc_match_sequence(self.h, s.ctypes.data if s.size else None, s.size, targetMatchLength, minMatchLength,
                                                     C.byref(n), off.ctypes.data, C.byref(no), ln.ctypes.data, C.byref(nl), ms))
        return s[:n.value].copy(), off[:no.value].copy(), ln[:nl.value].copy(), {"index_ms": ms[0], "query_ms": ms[1]}

    # -- row f4: the decoder's match-copy path (mbgccoder/MBGC_Decoder.cpp:243-675) ----------------------------------------------
    def decodeTargets(self, params, streams, refTotalLength, seqCounts, factors, g0IsTarget=True, g0SeqCount=0, rcInReference=True,
                      contigsIndividuallyReversed=True, lockPos=None, fetch=True):
        """streams: the six streams (bytes / uint8 arrays) in MBGCG_S_* order; seqCounts: contigs per target; factors:
        (unmatchedFractionFactor, unmatchedFractionRCFactor) per target or one pair for all.
        -> dict(contigs=[uint8 arrays], lens, unmatched, refPos, reached, batches, launches, parse_ms, copy_ms)"""
        keep = [_as_u8(x) for x in streams]
        inp = DecodeInput()
        for i, a in enumerate(keep):
            inp.stream[i] = a.ctypes.data if a.size else None
            inp.streamLen[i] = a.size
        n = len(seqCounts)
        cnt = (C.c_uint32 * max(n, 1))(*seqCounts)
        fac = np.asarray(factors, dtype=np.uint8).reshape(-1)
        if fac.size == 2:
            fac = np.tile(fac, max(n, 1))
        fac = np.ascontiguousarray(fac)
        inp.refTotalLength = refTotalLength
        inp.g0IsTarget = 1 if g0IsTarget else 0
        inp.g0SeqCount = g0SeqCount
        inp.rcInReference = 1 if rcInReference else 0
        inp.contigsIndividuallyReversed = 1 if contigsIndividuallyReversed else 0
        inp.nTargets = n
        inp.seqCounts = cnt
        inp.unmatchedFractionFactors = fac.ctypes.data_as(C.POINTER(C.c_uint8))
        lk = None
        if lockPos is not None:
            lk = (C.c_uint64 * max(n, 1))(*lockPos)
            inp.lockPos = lk
        res = DecodeResult()
        self._check(self.lib.mbgcg_decode_targets(self.h, C.byref(params), C.byref(inp), C.byref(res)))
        k = res.nContigs
        lens = np.ctypeslib.as_array(res.contigLen, shape=(k,)).copy() if k else np.zeros(0, dtype=np.uint64)
        unm = np.ctypeslib.as_array(res.contigUnmatched, shape=(k,)).copy() if k else np.zeros(0, dtype=np.uint64)
        out = dict(lens=lens, unmatched=unm, refPos=res.refPos, reached=res.reachedRefLengthCount, batches=res.batches,
                   launches=res.launches, parse_ms=res.parseMs, copy_ms=res.copyMs, dnaLen=res.dnaLen)
        if fetch:
            dna = np.empty(max(int(res.dnaLen), 1), dtype=np.uint8)
            self._check(self.lib.mbgcg_decode_fetch(self.h, 0, k, dna.ctypes.data))
            out["contigs"] = np.split(dna[:res.dnaLen], np.cumsum(lens)[:-1].astype(np.int64)) if k else []
        return out

    def decodeFetchRef(self, pos, n):
        out = np.empty(max(n, 1), dtype=np.uint8)
        self._check(self.lib.mbgcg_decode_fetch_ref(self.h, pos, n, out.ctypes.data))
        return out[:n]

    # -- row a14: the reference's legacy symbol packers (coders/SymbolsPackingFacility.cpp, coders/VarLenDNACoder.cpp) ------
    def packSymbols(self, symbols, src):
        """SymbolsPackingFacility(symbols).packSequence(src)"""
        a = _as_u8(src)
        out = np.empty(a.size + 8, dtype=np.uint8)
        n = C.c_uint64()
        self._check(self.lib.mbgcg_symbols_pack(self.h, bytes(symbols), len(symbols), a.ctypes.data, a.size, out.ctypes.data, C.byref(n)))
        return out[:n.value].copy()

    def unpackSymbols(self, symbols, packed, pos, length):
        """SymbolsPackingFacility(symbols).reverseSequence(packed, pos, length)"""
        p = _as_u8(packed)
        out = np.empty(max(length, 1), dtype=np.uint8)
        self._check(self.lib.mbgcg_symbols_unpack(self.h, bytes(symbols), len(symbols), p.ctypes.data, pos, length, out.ctypes.data))
        return out[:length].copy()

    def varLenEncode(self, book, src):
        """VarLenDNACoder(book).encode(src); book = the words separated by newlines"""
        a = _as_u8(src)
        out = np.empty(a.size + 8, dtype=np.uint8)
        n = C.c_uint64()
        self._check(self.lib.mbgcg_varlen_encode(self.h, book.encode("latin-1"), a.ctypes.data if a.size else None, a.size,
                                                 out.ctypes.data, out.size, C.byref(n)))
        return out[:n.value].copy()

    def varLenDecode(self, book, codes, expected_len):
        """VarLenDNACoder(book).decode(codes) -> exactly expected_len symbols"""
        c = _as_u8(codes)
        out = np.empty(max(expected_len, 1), dtype=np.uint8)
        self._check(self.lib.mbgcg_varlen_decode(self.h, book.encode("latin-1"), c.ctypes.data if c.size else None, c.size,
                                                 out.ctypes.data, expected_len))
        return out[:expected_len].copy()

    def deviceCount(self):
        return self.lib.mbgcg_device_count(self.h)

    def setHostMirror(self, on):
        self.lib.mbgcg_set_host_mirror(self.h, int(bool(on)))

    def compressSequential(self, images, cfg, loadedPosArr=None, pinned=True, on_event=None, keep_streams=True):
        """The reference's sequential (-t 1) target schedule over file images (bytes objects, in target order) through
        mbgcg_seq_*: MultipleGenomeMatchingProcessor::processTargetsWithParallelIO
        (/root/reference/matching/MultipleGenomeMatchingProcessor.cpp:232-313) with the encoder's hooks
        (mbgccoder/MBGC_Encoder.cpp:489-509) applied to the committed events.  Returns ({stream: bytes}, info) where the
        streams are what the encoder holds after the last target (the literal stream without the G0 bootstrap of
        MBGC_Encoder.cpp:34-37) and info has per-file meta data, headers, lock positions and the pass statistics."""
        keep = []

        def provider(i):
            a = _as_u8(images[i])
            p = self.lib.mbgcg_image_alloc(self.h, a.size) if (pinned and a.size) else None
            if p:
                C.memmove(p, a.ctypes.data, a.size)
                return p, a.size
            keep.append(a)
            return (a.ctypes.data if a.size else None), a.size

        return self._run_pass(len(images), provider, None, cfg, loadedPosArr, on_event, keep_streams)

    def compressFiles(self, paths, cfg, loadedPosArr=None, readers=None, on_event=None, keep_streams=True, pack=None):
        """compressSequential for files on disk: `readers` threads bring them (ahead of the pass, in order of need) into pinned
        memory of the handle's arena -- packed 4 bases per byte on the way (mbgcg_image_pack_file;
        pack=False or MBGC_B200_PACK_IMAGES=0: read as they are) -- or into pageable memory when the arena is full.
        gzip is not handled here (the reference's own reader does that in the drop-in)."""
        import threading
        if pack is None:
            pack = os.environ.get("MBGC_B200_PACK_IMAGES", "1") != "0"
        if on_event is None:
            # the whole thing inside the library: reader threads, pinned arena, the pass (no Python between a file and the device)
            n = len(paths)
            # all cores but three: the commit loop, the replica's worker and the stagers need some too (16 readers on 16 cores
            # take 40 % longer than 13)
            readers = max(1, min(readers or max(1, (os.cpu_count() or 1) - 3), 64, max(n, 1)))
            loaded = list(loadedPosArr) if loadedPosArr is not None else [self.getLoadingPosition()]
            arr = (C.c_uint64 * len(loaded))(*loaded)
            self._check(self.lib.mbgcg_seq_begin(self.h, C.byref(cfg), arr, len(loaded)))
            cpaths = (C.c_char_p * max(n, 1))(*[os.fsencode(x) for x in paths])
            host = (C.c_double * 2)()

            def call(res):
                return self.lib.mbgcg_seq_run_files(self.h, cpaths, n, readers, int(bool(pack)), int(bool(keep_streams)), C.byref(res), host)
            out, info = self._finish_run(call, keep_streams)
            info["host"] = dict(reader_busy_s=host[0] / 1e3, feed_wait_s=host[1] / 1e3, readers=readers, packed=bool(pack))
            return out, info
        pack = False   # the event loop pushes with mbgcg_seq_push_file
        n = len(paths)
        readers = max(1, min(readers or (os.cpu_count() or 1), 32, max(n, 1)))
        ready = [None] * n
        cond = threading.Condition()
        state = {"next": 0, "turn": 0, "consumed": 0, "stop": False}
        prof = {"reader_busy_s": 0.0, "reader_turn_wait_s": 0.0, "reader_alloc_wait_s": 0.0, "feed_wait_s": 0.0}
        arena_max = self.lib.mbgcg_image_max(self.h)
        keep = {}

        def work():
            while True:
                with cond:
                    i = state["next"]
                    if i >= n or state["stop"]:
                        return
                    state["next"] = i + 1
                    t0 = time.perf_counter()
                    cond.wait_for(lambda: i < state["consumed"] + 64 or state["stop"])
                    cond.wait_for(lambda: state["turn"] == i or state["stop"])   # pinned memory goes out in file order
                    prof["reader_turn_wait_s"] += time.perf_counter() - t0
                    if state["stop"]:
                        return
                size = os.path.getsize(paths[i])
                p = None
                want = self.lib.mbgcg_image_pack_bound(size) if pack else size
                t0 = time.perf_counter()
                if 0 < size and want <= arena_max:
                    for _ in range(20000):
                        p = self.lib.mbgcg_image_alloc(self.h, want)
                        if p or state["stop"]:
                            break
                        time.sleep(50e-6)
                t1 = time.perf_counter()
                with cond:
                    state["turn"] = i + 1
                    prof["reader_alloc_wait_s"] += t1 - t0
                    cond.notify_all()
                t0 = time.perf_counter()
                if p and pack:
                    used, fsize = C.c_uint64(), C.c_uint64()
                    rc = self.lib.mbgcg_image_pack_file(os.fsencode(paths[i]), p, want, C.byref(used), C.byref(fsize))
                    if rc == 0 and fsize.value != size:
                        raise IOError(f"{paths[i]} changed size while it was read")
                    if rc == 0:
                        with cond:
                            ready[i] = (p, used.value | (1 << 63))
                            prof["reader_busy_s"] += time.perf_counter() - t0
                            cond.notify_all()
                        continue
                    self.lib.mbgcg_image_free(self.h, p)   # too many exceptional bytes: the file travels as it is
                    p = None
                if p:
                    buf = (C.c_char * size).from_address(p)
                else:
                    arr = np.empty(size, dtype=np.uint8)
                    keep[i] = arr
                    buf = memoryview(arr)
                    p = arr.ctypes.data if size else None
                with open(paths[i], "rb", buffering=0) as f:
                    got = 0
                    mv = memoryview(buf).cast("B")
                    while got < size:
                        k = f.readinto(mv[got:])
                        if not k:
                            raise IOError(f"short read on {paths[i]}")
                        got += k
                with cond:
                    ready[i] = (p, size)
                    prof["reader_busy_s"] += time.perf_counter() - t0
                    cond.notify_all()

        threads = [threading.Thread(target=work, daemon=True) for _ in range(readers)]
        for t in threads:
            t.start()

        def provider(i, block):
            with cond:
                if ready[i] is None:
                    if not block:
                        return None
                    t0 = time.perf_counter()
                    cond.wait_for(lambda: ready[i] is not None)
                    prof["feed_wait_s"] += time.perf_counter() - t0
                return ready[i]

        def consumed(i, drop=True):
            with cond:
                state["consumed"] = max(state["consumed"], i + 1)
                if drop:   # only once the file's FILE_END has been seen: a pageable image must stay valid until then
                    keep.pop(i, None)
                cond.notify_all()

        try:
            out, info = self._run_pass(n, None, (provider, consumed), cfg, loadedPosArr, on_event, keep_streams)
            info["host"] = dict(prof, readers=readers, packed=bool(pack))
            return out, info
        finally:
            with cond:
                state["stop"] = True
                cond.notify_all()
            for t in threads:
                t.join()

    def _run_pass(self, n, provider, async_provider, cfg, loadedPosArr, on_event, keep_streams):
        loaded = list(loadedPosArr) if loadedPosArr is not None else [self.getLoadingPosition()]
        arr = (C.c_uint64 * len(loaded))(*loaded)
        self._check(self.lib.mbgcg_seq_begin(self.h, C.byref(cfg), arr, len(loaded)))
        if on_event is None:
            return self._run_pass_to_end(n, provider, async_provider, keep_streams)
        return self._run_pass_events(n, provider, async_provider, cfg, on_event, keep_streams)

    def _run_pass_to_end(self, n, provider, async_provider, keep_streams):
        """mbgcg_seq_run: the whole pass inside the library, images pulled through a callback"""
        state = {"next": 0, "err": None}

        def next_image(user, block, bytes_out, n_out):
            try:
                i = state["next"]
                if i >= n:
                    return -1
                if provider is not None:
                    p, size = provider(i)
                else:
                    got = async_provider[0](i, bool(block))
                    if got is None:
                        return 0
                    p, size = got
                    async_provider[1](i - 1, False)   # read-ahead window follows the pushes; pageable images stay until the end
                bytes_out[0] = p
                n_out[0] = size
                state["next"] = i + 1
                return 1
            except BaseException as e:  # noqa: BLE001 -- must not propagate through the C frame
                state["err"] = e
                return -1

        cb = NEXT_IMAGE_FN(next_image)

        def call(res):
            rc = self.lib.mbgcg_seq_run(self.h, cb, None, int(bool(keep_streams)), C.byref(res))
            if state["err"] is not None:
                raise state["err"]
            return rc
        return self._finish_run(call, keep_streams)

    def _finish_run(self, call, keep_streams):
        """call(res) runs the pass to its end (mbgcg_seq_run / mbgcg_seq_run_files); -> (streams, info), pass closed"""
        res = SeqResult()
        st = SeqStats()
        try:
            self._check(call(res))
            nf = int(res.nFiles)

            def arr64(p):
                return np.ctypeslib.as_array(C.cast(p, u64p), shape=(nf,)).astype(np.int64).tolist() if nf else []
            out = {name: (C.string_at(res.stream[s], res.streamLen[s]) if (keep_streams and res.streamLen[s]) else b"")
                   for s, name in enumerate(STREAM_NAMES)}
            hdrs = C.string_at(res.headers, res.headersLen).split(b"\n")[:-1] if res.headersLen else []
            counts = arr64(res.seqCounts)
            files_of = [f for f, c in enumerate(counts) for _ in range(c)]
            info = dict(headers=list(zip(files_of, hdrs)), seqCounts=counts, dnaLineLens=arr64(res.dnaLineLens),
                        refExtSizes=arr64(res.refExtSizes), lockPos=[int(x) & (2**64 - 1) for x in arr64(res.lockPos)],
                        unmatched=int(res.unmatched), matches=int(res.matches), extensions=int(res.extensions),
                        status=(np.ctypeslib.as_array(C.cast(res.status, C.POINTER(C.c_int32)), shape=(nf,)).tolist() if nf else []),
                        contigs=int(res.nContigs), digest=int(res.digest),
                        stream_bytes={name: int(res.streamLen[s]) for s, name in enumerate(STREAM_NAMES)})
        finally:
            self.lib.mbgcg_seq_end(self.h, C.byref(st))
        info["stats"] = {name: getattr(st, name) for name, _ in SeqStats._fields_}
        return out, info

    def _run_pass_events(self, n, provider, async_provider, cfg, on_event, keep_streams):
        out = {name: [] for name in STREAM_NAMES}
        sizes = {name: 0 for name in STREAM_NAMES}
        info = dict(headers=[], seqCounts=[], dnaLineLens=[], refExtSizes=[], lockPos=[], unmatched=0, matches=0,
                    extensions=0, status=[], contigs=0)
        nxt = 0
        finished = False
        ev = SeqEvent()
        try:
            while True:
                while nxt < n and self.lib.mbgcg_seq_wants_input(self.h):
                    if provider is not None:
                        p, size = provider(nxt)
                    else:
                        got = async_provider[0](nxt, False)
                        if got is None:
                            break
                        p, size = got
                    self._check(self.lib.mbgcg_seq_push_file(self.h, p, size))
                    nxt += 1
                if nxt == n and not finished:
                    self._check(self.lib.mbgcg_seq_finish_input(self.h))
                    finished = True
                self._check(self.lib.mbgcg_seq_next(self.h, C.byref(ev)))
                if on_event is not None:
                    on_event(ev)
                if ev.type == EV_DONE:
                    break
                if ev.type == EV_NEED_INPUT:
                    if finished:
                        raise MatcherError(-2, "the pass asks for input after finish_input")
                    if provider is not None:
                        p, size = provider(nxt)
                    else:
                        p, size = async_provider[0](nxt, True)
                    self._check(self.lib.mbgcg_seq_push_file(self.h, p, size))
                    nxt += 1
                    continue
                if ev.type == EV_CONTIG:
                    for s, name in enumerate(STREAM_NAMES):
                        if ev.segLen[s]:
                            sizes[name] += int(ev.segLen[s])
                            if keep_streams:
                                out[name].append(C.string_at(ev.seg[s], ev.segLen[s]))
                    sizes["literals"] += 1
                    if keep_streams:
                        out["literals"].append(bytes([SEQ_SEPARATOR_MARK]))          # processAfterSequence
                        info["headers"].append((int(ev.file), C.string_at(ev.hdr, ev.hdrLen) if ev.hdrLen else b""))
                    info["unmatched"] += int(ev.unmatched)
                    info["matches"] += int(ev.nMatches)
                    info["extensions"] += int(ev.extended)
                    info["contigs"] += 1
                elif ev.type == EV_FILE_END:
                    if cfg.pack.enableExtensionsWithMismatches:                   # processAfterTarget
                        sizes["gapFlags"] += 1
                        if keep_streams:
                            out["gapFlags"].append(bytes([FILE_SEPARATOR_MARK]))
                    info["seqCounts"].append(int(ev.nContigs))
                    info["dnaLineLens"].append(int(ev.dnaLineLen))
                    info["refExtSizes"].append(int(ev.refExtSize))
                    info["lockPos"].append(int(ev.lockPos))
                    info["status"].append(int(ev.status))
                    if async_provider is not None:
                        async_provider[1](int(ev.file))
        finally:
            st = SeqStats()
            self.lib.mbgcg_seq_end(self.h, C.byref(st))
        info["stats"] = {name: getattr(st, name) for name, _ in SeqStats._fields_}
        info["stream_bytes"] = sizes
        return {name: b"".join(v) for name, v in out.items()}, info

    def close(self):
        if getattr(self, "h", None):
            self.lib.mbgcg_destroy(self.h)
            self.h = None

    def __del__(self):
        self.close()

    def _check(self, rc):
        if rc != 0:
            raise MatcherError(rc, self.lib.mbgcg_last_error(self.h).decode())

    # -- reference-named methods -------------------------------------------------------------------
    def disableSlidingWindow(self):
        self.lib.mbgcg_disable_sliding_window(self.h)

    def disableCircularBuffer(self):
        self.lib.mbgcg_disable_circular_buffer(self.h)

    def setSlidingWindowSize(self, factor):
        self.lib.mbgcg_set_sliding_window_size(self.h, factor)

    def loadRef(self, refText, loadRCRef=False, addRegionSeparators=True, regionSeparator=0):
        a = _as_u8(refText)
        self._check(self.lib.mbgcg_load_ref(self.h, a.ctypes.data, a.size, int(loadRCRef), int(addRegionSeparators),
                                            bytes([regionSeparator])))

    def loadSeparator(self, regionSeparator=0):
        self._check(self.lib.mbgcg_load_separator(self.h, bytes([regionSeparator])))

    def matchTexts(self, destText, minMatchLength=32, matchingLockPos=NO_LOCK):
        a = _as_u8(destText)
        out = C.c_void_p()
        n = C.c_uint64()
        self._check(self.lib.mbgcg_match(self.h, a.ctypes.data, a.size, minMatchLength, matchingLockPos,
                                         C.byref(out), C.byref(n)))
        return self._matches(out, n.value)

    def matchBatch(self, texts, minMatchLength=32, lockPos=None, copy=True, fetch=True, split=True):
        """list of texts -> list of (n_i, 3) arrays, all matched against the same frozen state.
        copy=False returns views of the library's pinned result buffer (valid until the next match call)."""
        arrs = [_as_u8(t) for t in texts]
        n = len(arrs)
        ptrs = (C.c_void_p * n)(*[a.ctypes.data for a in arrs])
        lens = (C.c_uint64 * n)(*[a.size for a in arrs])
        locks = (C.c_uint64 * n)(*lockPos) if lockPos is not None else None
        out, counts = C.c_void_p(), C.c_void_p()
        self._check(self.lib.mbgcg_match_batch(self.h, n, ptrs, lens, minMatchLength, locks,
                                               C.byref(out) if fetch else None, C.byref(counts)))
        if not fetch:  # lists stay on the device for packBatch(matches=None)
            return np.ctypeslib.as_array(C.cast(counts, u64p), shape=(n,)).copy()
        return self._split(out, counts, n, copy, split)

    def matchBatchDevice(self, dTextPtr, offsets, lens, minMatchLength=32, lockPos=None, copy=True, fetch=True,
                         split=True):
        """texts already resident in device memory (see mbgcg_match_batch_device for the layout rules).
        fetch=False leaves the lists on the device for a packBatch(matches=None) call and returns the counts.
        split=False returns (all lists concatenated, counts per text) instead of one array per text."""
        n = len(lens)
        offs, ls = _u64_array(offsets), _u64_array(lens)   # uint64 numpy arrays pass through without a copy
        locks = _u64_array(lockPos) if lockPos is not None else None
        out, counts = C.c_void_p(), C.c_void_p()
        self._check(self.lib.mbgcg_match_batch_device(self.h, dTextPtr, n, offs, ls, minMatchLength, locks,
                                                      C.byref(out) if fetch else None, C.byref(counts)))
        if not fetch:
            return np.ctypeslib.as_array(C.cast(counts, u64p), shape=(n,)).copy()
        return self._split(out, counts, n, copy, split)

    def packBatch(self, params, texts, matches=None, lockPos=None, loadedPosArr=(), lens=None, raw=False):
        """MBGC_Encoder::processMatches for a list of texts (mbgcg_pack_batch).  matches: list of (n_i, 3) uint64
        arrays (posSrcText, length, posDestText), or None to pack the lists of the immediately preceding match call
        (still on the device; texts may then be None with lens given).  Returns a dict: per stream name a list of
        bytes per text, plus "unmatched" and "kept" (arrays per text).  raw=True skips the per-text slicing and
        returns {"stream_bytes": {name: total}, "unmatched", "kept"} (the streams stay in the library's pinned buffers)."""
        if texts is not None:
            arrs = [_as_u8(t) for t in texts]
            n = len(arrs)
            ptrs = (C.c_void_p * n)(*[a.ctypes.data for a in arrs])
            ls = (C.c_uint64 * n)(*[a.size for a in arrs])
        else:
            n = len(lens)
            ptrs = None
            ls = _u64_array(lens)
        if matches is not None:
            cnt = (C.c_uint64 * n)(*[len(m) for m in matches])
            tot = sum(len(m) for m in matches)
            flat = np.zeros((max(tot, 1), 4), dtype=np.uint64)
            at = 0
            for m in matches:
                if len(m):
                    flat[at:at + len(m), :3] = np.asarray(m, dtype=np.uint64).reshape(-1, 3)
                    at += len(m)
            mptr = flat.ctypes.data
        else:
            cnt, mptr = None, None
        locks = (C.c_uint64 * n)(*lockPos) if lockPos is not None else None
        nl = len(loadedPosArr)
        loaded = (C.c_uint64 * max(nl, 1))(*loadedPosArr)
        res = _PackResult()
        self._check(self.lib.mbgcg_pack_batch(self.h, C.byref(params), n, ptrs, ls, mptr, cnt, locks, loaded, nl,
                                              C.byref(res)))
        out = {}
        if raw:
            out["stream_bytes"] = {name: int(C.cast(res.streamOff[s], u64p)[n]) if n else 0
                                   for s, name in enumerate(STREAM_NAMES)}
        for s, name in enumerate(STREAM_NAMES if not raw else ()):
            off = np.ctypeslib.as_array(C.cast(res.streamOff[s], u64p), shape=(n + 1,)).copy() if n else np.zeros(1, np.uint64)
            total = int(off[-1])
            data = bytes(np.ctypeslib.as_array(C.cast(res.stream[s], u8p), shape=(total,))) if total else b""
            out[name] = [data[int(off[i]):int(off[i + 1])] for i in range(n)]
        out["unmatched"] = np.ctypeslib.as_array(C.cast(res.unmatched, u64p), shape=(n,)).copy() if n else np.zeros(0, np.uint64)
        out["kept"] = np.ctypeslib.as_array(C.cast(res.kept, u64p), shape=(n,)).copy() if n else np.zeros(0, np.uint64)
        return out

    def packKeptMatches(self):
        """textMatches as pass 1 of the last packBatch left them (all texts concatenated), shape (n, 3)"""
        out = C.c_void_p()
        n = C.c_uint64()
        self._check(self.lib.mbgcg_pack_kept_matches(self.h, C.byref(out), C.byref(n)))
        return self._matches(out, n.value)

    def setLiteralRefExt(self, refLiteralMinimalLengthExt=None, refLiteralBeforeAfterExt=0, unmatchedFractionFactor=128):
        """MBGC_Encoder::processLiteral on (the reference's developer option -N / -F) or, with None, off again
        (mbgcg_set_literal_ref_ext): every later packBatch also lists the literal runs that extend the reference."""
        m = 0xFFFFFFFFFFFFFFFF if refLiteralMinimalLengthExt is None else int(refLiteralMinimalLengthExt)
        self._check(self.lib.mbgcg_set_literal_ref_ext(self.h, m, int(refLiteralBeforeAfterExt), int(unmatchedFractionFactor)))

    def packLiteralRefExt(self):
        """what processLiteral appended to targetRefExtensions during the last packBatch (mbgcg_pack_literal_ref_ext):
        (n, 3) uint32 array of (text index, start, length), in append order"""
        out = C.c_void_p()
        n = C.c_uint64()
        self._check(self.lib.mbgcg_pack_literal_ref_ext(self.h, C.byref(out), C.byref(n)))
        if not n.value:
            return np.zeros((0, 3), dtype=np.uint32)
        a = np.ctypeslib.as_array(C.cast(out, C.POINTER(C.c_uint32)), shape=(n.value, 4))
        return a[:, :3].copy()

    def packTiming(self):
        ms = (C.c_float * 4)()
        launches = C.c_uint64()
        sweeps = (C.c_int * 2)()
        self.lib.mbgcg_pack_timing(self.h, ms, C.byref(launches), sweeps)
        return {"h2d": ms[0], "kernels": ms[1], "d2h": ms[2], "total": ms[3], "launches": launches.value,
                "sweeps": (sweeps[0], sweeps[1])}

    def ingestFasta(self, data):
        """kseq_read_lossless_fasta (/root/reference/utils/kseq.h:233-274) over a whole file image on the device
        (mbgcg_ingest_fasta).  Returns a dict: records = (n, 4) uint64 array (hdrOff, hdrLen, textOff, seqLen), status
        (-1 end of file, -3 not FASTA, -4 DNA not well-formed), dnaLineLen, dTexts (device pointer of the sequences: pass
        it with textOff / seqLen to matchBatchDevice)."""
        a = _as_u8(data)
        recs, n, dtexts = C.c_void_p(), C.c_uint64(), C.c_void_p()
        dna, status = C.c_uint64(), C.c_int()
        self._check(self.lib.mbgcg_ingest_fasta(self.h, a.ctypes.data if a.size else None, a.size, C.byref(recs), C.byref(n),
                                                C.byref(dtexts), C.byref(dna), C.byref(status)))
        r = (np.ctypeslib.as_array(C.cast(recs, u64p), shape=(n.value, 4)).copy() if n.value
             else np.zeros((0, 4), dtype=np.uint64))
        return {"records": r, "status": status.value, "dnaLineLen": dna.value, "dTexts": dtexts.value}

    def ingestTiming(self):
        ms = (C.c_float * 3)()
        launches = C.c_uint64()
        self.lib.mbgcg_ingest_timing(self.h, ms, C.byref(launches))
        return {"h2d": ms[0], "kernels": ms[1], "total": ms[2], "launches": launches.value}

    def copyIngested(self, record, seqLen):
        out = np.empty(int(seqLen), dtype=np.uint8)
        if seqLen:
            self._check(self.lib.mbgcg_copy_ingested(self.h, record, out.ctypes.data))
        return out

    def acquireWorkerMatchingLockPos(self):
        return self.lib.mbgcg_acquire_lock(self.h)

    def releaseWorkerMatchingLockPos(self, lockValue):
        self._check(self.lib.mbgcg_release_lock(self.h, lockValue))

    def getRef(self, n=None):
        n = self.getMaxRefLength() if n is None else n
        return np.ctypeslib.as_array(C.cast(self.lib.mbgcg_host_ref(self.h), u8p), shape=(n,))

    def getMaxRefLength(self):
        return self.lib.mbgcg_max_ref_length(self.h)

    def getRefLength(self):
        return self.lib.mbgcg_ref_length(self.h)

    def getLoadingPosition(self):
        return self.lib.mbgcg_loading_position(self.h)

    def getLoadedRefLength(self):
        return self.lib.mbgcg_loaded_ref_length(self.h)

    def setPosition(self, refPos, reachedRefLengthCount):
        self.lib.mbgcg_set_position(self.h, refPos, reachedRefLengthCount)

    # -- inspection ----------------------------------------------------------------------------------
    def state(self):
        s = _State()
        self.lib.mbgcg_get_state(self.h, C.byref(s))
        return {n: getattr(s, n) for n, _ in _State._fields_}

    def table(self):
        n = self.state()["hashSize"]
        out = np.empty(n, dtype=np.uint32)
        self._check(self.lib.mbgcg_copy_table(self.h, out.ctypes.data, n))
        return out

    def deviceRef(self, start, n):
        out = np.empty(n, dtype=np.uint8)
        self._check(self.lib.mbgcg_copy_device_ref(self.h, start, n, out.ctypes.data))
        return out

    def stagedText(self, index, n):
        """text `index` of the last host-buffer batch as it arrived on the device (n = its length)"""
        out = np.empty(n, dtype=np.uint8)
        self._check(self.lib.mbgcg_copy_staged_text(self.h, index, out.ctypes.data))
        return out

    def setStageTiming(self, on):
        """switch the per-stage CUDA events of match calls off / on (timing() returns zeros while off)"""
        self.lib.mbgcg_set_stage_timing(self.h, int(bool(on)))

    def timing(self):
        ms = (C.c_float * 12)()
        self.lib.mbgcg_last_timing(self.h, ms)
        return {name: ms[i] for i, name in enumerate(T_NAMES)}

    def counters(self):
        c = (C.c_uint64 * 8)()
        self.lib.mbgcg_last_counters(self.h, c)
        return {"positions": c[0], "seeds": c[1], "runs": c[2], "matches": c[3], "launches": c[4],
                "passes": c[5], "chunks": c[6], "lanes_full_probe": c[7]}

    def wireBytes(self):
        """bytes the last match call moved over PCIe: (host -> device, device -> host)"""
        w = (C.c_uint64 * 2)()
        self.lib.mbgcg_last_wire_bytes(self.h, w)
        return int(w[0]), int(w[1])

    def loadRefTiming(self):
        """last loadRef: device ms (CUDA events), host wall ms, reference bytes appended"""
        ms = (C.c_float * 2)()
        n = C.c_uint64()
        self.lib.mbgcg_load_ref_timing(self.h, ms, C.byref(n))
        return {"device_ms": ms[0], "wall_ms": ms[1], "appended": int(n.value)}

    def stream(self):
        """cudaStream_t (as int) all work of this matcher is issued on"""
        return self.lib.mbgcg_stream(self.h)

    def totalLaunches(self):
        return self.lib.mbgcg_total_launches(self.h)

    # -- helpers -------------------------------------------------------------------------------------
    @staticmethod
    def _matches(out, n, copy=True):
        if n == 0:
            return np.zeros((0, 3), dtype=np.uint64)
        a = np.ctypeslib.as_array(C.cast(out, u64p), shape=(n, 4))
        return a[:, :3].copy() if copy else a[:, :3]

    def _split(self, out, counts, n, copy=True, split=True):
        cnt = np.ctypeslib.as_array(C.cast(counts, u64p), shape=(n,)).copy() if n else np.zeros(0, np.uint64)
        allm = self._matches(out, int(cnt.sum()), copy)
        if not split:
            return allm, cnt
        ends = np.cumsum(cnt.astype(np.int64))
        return [allm[int(e - c):int(e)] for c, e in zip(cnt, ends)]
