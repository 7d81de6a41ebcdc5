"""Shared pieces of the decoder tests (row f4): bindings of the oracle's decoder restatement (oracle/mbgc_oracle_dec.c) and
of the reference's own MBGC_Decoder<false> behind oracle/ref_decoder_driver.cpp, and an in-memory emulation of the
reference's deterministic (-t 1) encoder schedule that yields the six streams plus what the archive's side streams would
tell a decoder (contigs per target, extension factors, the final reference length).  Test infrastructure."""
import ctypes as C
import os

import numpy as np

from util import NO_LOCK, ORACLE_DIR, REF_DIR, OracleMatcher, PackParams, oracle_lib, synth_base, synth_mutant, u64p, u8p

STREAMS = ("literals", "mapOff", "mapOff5", "mapLen", "gapDelta", "gapFlags")
SEQ_SEPARATOR_MARK = (ord('"') + 128) & 0xFF


class DecStreams(C.Structure):
    _fields_ = [(n, t) for name in STREAMS for n, t in ((name, u8p), ("n_" + name, C.c_uint64))]


_dec = None


def dec_oracle():
    global _dec
    if _dec is None:
        L = oracle_lib()
        src = os.path.join(ORACLE_DIR, "mbgc_oracle_dec.c")
        if os.path.getmtime(os.path.join(ORACLE_DIR, "libmbgc_oracle.so")) < os.path.getmtime(src):
            raise RuntimeError("oracle/libmbgc_oracle.so is older than mbgc_oracle_dec.c: run make -C oracle oracle")
        L.mbo_dec_create.restype = C.c_void_p
        L.mbo_dec_create.argtypes = [C.POINTER(PackParams), C.c_uint64, C.c_int, C.c_int, C.POINTER(DecStreams)]
        L.mbo_dec_destroy.argtypes = [C.c_void_p]
        L.mbo_code2mismatch.restype = C.c_uint8
        L.mbo_code2mismatch.argtypes = [C.c_uint8, C.c_uint8]
        L.mbo_dec_init_reference.argtypes = [C.c_void_p]
        L.mbo_dec_decode_reference.restype = C.c_uint64
        L.mbo_dec_decode_reference.argtypes = [C.c_void_p, C.c_uint32, u8p, u64p]
        L.mbo_dec_decode_sequence.restype = C.c_uint64
        L.mbo_dec_decode_sequence.argtypes = [C.c_void_p, C.c_uint64, u8p, u64p]
        L.mbo_dec_load_ref.argtypes = [C.c_void_p, u8p, C.c_uint64, C.c_uint64, C.c_int]
        L.mbo_dec_decode_target.restype = C.c_uint64
        L.mbo_dec_decode_target.argtypes = [C.c_void_p, C.c_uint32, C.c_int, C.c_int, C.c_uint64, u8p, u64p, u64p]
        L.mbo_dec_ref.restype = C.c_void_p
        L.mbo_dec_ref.argtypes = [C.c_void_p]
        L.mbo_dec_ref_pos.restype = C.c_uint64
        L.mbo_dec_ref_pos.argtypes = [C.c_void_p]
        L.mbo_dec_reached_count.restype = C.c_int
        L.mbo_dec_reached_count.argtypes = [C.c_void_p]
        L.mbo_dec_cursors.argtypes = [C.c_void_p, u64p]
        _dec = L
    return _dec


_refd = None


def ref_decoder():
    """the reference's own decoder class, or None where oracle/_ref does not travel"""
    global _refd
    if _refd is None:
        path = os.path.join(REF_DIR, "libref_decoder.so")
        if not os.path.exists(path):
            return None
        L = C.CDLL(path)
        L.refd_create.restype = C.c_void_p
        L.refd_create.argtypes = [C.POINTER(C.c_int), C.c_uint64] + [u8p, C.c_uint64] * 6 + [C.POINTER(C.c_uint32), u8p, u64p, C.c_uint32]
        L.refd_destroy.argtypes = [C.c_void_p]
        L.refd_init_reference.argtypes = [C.c_void_p]
        L.refd_decode_reference.argtypes = [C.c_void_p]
        L.refd_decode_target.restype = C.c_uint64
        L.refd_decode_target.argtypes = [C.c_void_p, C.c_uint32, u8p, C.c_uint64]
        L.refd_decode_sequence.restype = C.c_uint64
        L.refd_decode_sequence.argtypes = [C.c_void_p, C.c_uint64, u8p, C.c_uint64, u64p]
        L.refd_load_ref.argtypes = [C.c_void_p, u8p, C.c_uint64, C.c_uint64, C.c_int]
        L.refd_ref.restype = C.c_void_p
        L.refd_ref.argtypes = [C.c_void_p]
        L.refd_ref_pos.restype = C.c_uint64
        L.refd_ref_pos.argtypes = [C.c_void_p]
        L.refd_reached_count.restype = C.c_int
        L.refd_reached_count.argtypes = [C.c_void_p]
        L.refd_set_ref_pos.argtypes = [C.c_void_p, C.c_uint64]
        L.refd_cursors.argtypes = [C.c_void_p, u64p]
        _refd = L
    return _refd


def _ptr(b):
    a = np.frombuffer(b, dtype=np.uint8) if len(b) else np.zeros(8, dtype=np.uint8)
    return a, a.ctypes.data_as(u8p)


class Encoded:
    """streams + side information of one encoded collection"""

    def __init__(self, streams, seq_counts, factors, ref_total, g0_is_target, contigs, pp, rc_in_ref=True):
        self.streams = streams          # name -> bytes
        self.seq_counts = seq_counts    # contigs per file, file 0 first
        self.factors = factors          # (unmatchedFractionFactor, unmatchedFractionRCFactor)
        self.ref_total = ref_total      # the archive's refTotalLength
        self.g0_is_target = g0_is_target
        self.contigs = contigs          # expected output: list (per file) of list of uint8 arrays
        self.pp = pp
        self.rc_in_ref = rc_in_ref

    @property
    def targets(self):
        """(file index, contig count) of every target in decode order"""
        first = 0 if self.g0_is_target else 1
        return [(f, self.seq_counts[f]) for f in range(first, len(self.seq_counts))]


def encode_collection(files, pp=None, rc_factor=8, skip_margin=16, g0_is_target=True, limit=None):
    """files: list (per file) of list of uint8 arrays (contigs).  The reference's sequential schedule
    (MultipleGenomeMatchingProcessor.cpp:232-313; tests/seq_schedule.py does the same from FASTA files): per contig
    matchTexts, processMatches, extension decision, loadRef; per file separator + refExtLoadedPosArr."""
    lib = oracle_lib()
    pp = pp or PackParams.for_mode(1)
    total = sum(int(c.size) for f in files for c in f)
    if limit is None:
        limit = max(1 << 22, 4 * total + (1 << 16))
    m = OracleMatcher(limit, 32, 16, 1, skip_margin)
    m.disable_sliding_window()
    streams = lib.mbo_streams_create()
    if g0_is_target:
        c0 = files[0][0]
        head = bytes(c0) + bytes([SEQ_SEPARATOR_MARK])
        m.load_ref(c0, rc=True, add_sep=True)
        first = 0
    else:
        head = b"".join(bytes(c) + bytes([SEQ_SEPARATOR_MARK]) for c in files[0])
        m.load_ref(np.concatenate(files[0]), rc=True, add_sep=True)
        first = 1
    loaded = [m.loading_position]
    for contigs in files[first:]:
        start = m.loaded_ref_length
        for a in contigs:
            a = np.ascontiguousarray(a, dtype=np.uint8)
            mt = np.ascontiguousarray(m.match(a, 32))
            n = C.c_uint64(len(mt))
            buf = mt if len(mt) else np.zeros((1, 3), dtype=np.uint64)
            arr = (C.c_uint64 * len(loaded))(*loaded)
            unmatched = lib.mbo_process_matches(streams, C.byref(pp), m.h, buf.ctypes.data_as(C.c_void_p), C.byref(n),
                                                a.ctypes.data_as(u8p), a.size, NO_LOCK, arr, len(loaded))
            if unmatched * pp.unmatchedFractionFactor > a.size:
                m.load_ref(a, rc=unmatched * rc_factor > a.size, add_sep=True)
            lib.mbo_after_sequence(streams)
        lib.mbo_after_target(streams, C.byref(pp))
        if pp.lazyDecompressionSupport:
            m.load_separator(0)
        loaded.append(loaded[-1] + (m.loaded_ref_length - start))
    out = {}
    for i, name in enumerate(STREAMS):
        k = lib.mbo_stream_size(streams, i)
        p = lib.mbo_stream_data(streams, i)
        out[name] = bytes(np.ctypeslib.as_array(C.cast(p, u8p), shape=(k,))) if k else b""
    out["literals"] = head + out["literals"]
    lib.mbo_streams_destroy(streams)
    ref_total = m.ref_length
    m.close()
    return Encoded(out, [len(f) for f in files], (pp.unmatchedFractionFactor, rc_factor), ref_total, g0_is_target,
                   [[np.ascontiguousarray(c, dtype=np.uint8) for c in f] for f in files], pp)


def synth_files(seed, n_files, contigs_per_file, length, p_snp, p_indel):
    """a seeded collection: every file holds mutants of pieces of one base genome"""
    base = synth_base(length * contigs_per_file, seed)
    files = []
    for f in range(n_files):
        g = base if f == 0 else synth_mutant(base, seed * 7919 + f, p_snp, p_indel)
        cuts = np.linspace(0, g.size, contigs_per_file + 1).astype(np.int64)
        files.append([np.ascontiguousarray(g[cuts[i]:cuts[i + 1]]) for i in range(contigs_per_file)])
    return files


def _streams_struct(enc):
    keep = []
    s = DecStreams()
    for name in STREAMS:
        a, p = _ptr(enc.streams[name])
        keep.append(a)
        setattr(s, name, p)
        setattr(s, "n_" + name, len(enc.streams[name]))
    return s, keep


def oracle_decode(enc, locks=None, contigs_individually_reversed=True):
    """-> (per target: list of contigs, unmatched counts), final reference bytes, refPos, cursors"""
    L = dec_oracle()
    s, keep = _streams_struct(enc)
    d = L.mbo_dec_create(C.byref(enc.pp), enc.ref_total, 1 if enc.rc_in_ref else 0, 1 if contigs_individually_reversed else 0, C.byref(s))
    try:
        if enc.g0_is_target:
            L.mbo_dec_init_reference(d)
            g0 = None
        else:
            cap = sum(int(c.size) for c in enc.contigs[0]) + 16
            out = np.zeros(cap, dtype=np.uint8)
            lens = (C.c_uint64 * enc.seq_counts[0])()
            n = L.mbo_dec_decode_reference(d, enc.seq_counts[0], out.ctypes.data_as(u8p), lens)
            g0 = np.split(out[:n], np.cumsum(list(lens))[:-1])
        res = []
        for t, (f, cnt) in enumerate(enc.targets):
            cap = 2 * sum(int(c.size) for c in enc.contigs[f]) + (1 << 16)
            out = np.zeros(cap, dtype=np.uint8)
            lens = (C.c_uint64 * max(cnt, 1))()
            unm = (C.c_uint64 * max(cnt, 1))()
            lock = enc.ref_total if locks is None else locks[t]
            n = L.mbo_dec_decode_target(d, cnt, enc.factors[0], enc.factors[1], lock, out.ctypes.data_as(u8p), lens, unm)
            parts = np.split(out[:n].copy(), np.cumsum(list(lens)[:cnt])[:-1]) if cnt else []
            res.append((parts, list(unm)[:cnt]))
        ref = np.ctypeslib.as_array(C.cast(L.mbo_dec_ref(d), u8p), shape=(enc.ref_total,)).copy()
        cur = (C.c_uint64 * 6)()
        L.mbo_dec_cursors(d, cur)
        return res, g0, ref, L.mbo_dec_ref_pos(d), list(cur)
    finally:
        L.mbo_dec_destroy(d)


def dec_params_array(pp, rc_in_ref=True, contigs_individually_reversed=True):
    return (C.c_int * 11)(pp.enableExtensionsWithMismatches, pp.gapDepthOffsetEncoding, pp.gapDepthMismatchesEncoding,
                          pp.lazyDecompressionSupport, pp.frugal64bitLenEncoding, pp.mmsMismatchesInitialScore,
                          pp.mmsMismatchesScoreThreshold, pp.mmsMismatchPenalty, pp.mmsMatchBonus, 1 if rc_in_ref else 0,
                          1 if contigs_individually_reversed else 0)


def reference_decode(R, enc, locks=None, contigs_individually_reversed=True):
    """the same through the reference's own MBGC_Decoder<false>::decodeTarget"""
    keep = []
    args = []
    for name in STREAMS:
        a, p = _ptr(enc.streams[name])
        keep.append(a)
        args += [p, len(enc.streams[name])]
    tg = enc.targets
    counts = (C.c_uint32 * (len(tg) + 1))(enc.seq_counts[0], *[c for _, c in tg])
    fac = np.array([enc.factors[0], enc.factors[1]] * max(len(tg), 1), dtype=np.uint8)
    lk = (C.c_uint64 * len(tg))(*locks) if locks is not None else None
    h = R.refd_create(dec_params_array(enc.pp, enc.rc_in_ref, contigs_individually_reversed), enc.ref_total, *args, counts,
                      fac.ctypes.data_as(u8p), lk, len(tg))
    try:
        if enc.g0_is_target:
            R.refd_init_reference(h)
        else:
            R.refd_decode_reference(h)
        res = []
        for t, (f, cnt) in enumerate(tg):
            cap = 2 * sum(int(c.size) for c in enc.contigs[f]) + (1 << 16)
            out = np.zeros(cap, dtype=np.uint8)
            n = R.refd_decode_target(h, t, out.ctypes.data_as(u8p), cap)
            assert n <= cap
            lines = bytes(out[:n]).split(b"\n")
            assert lines[-1] == b"" and all(x == b">h" for x in lines[0:-1:2])
            res.append([np.frombuffer(x, dtype=np.uint8) for x in lines[1:-1:2]])
        ref = np.ctypeslib.as_array(C.cast(R.refd_ref(h), u8p), shape=(enc.ref_total,)).copy()
        cur = (C.c_uint64 * 6)()
        R.refd_cursors(h, cur)
        return res, ref, R.refd_ref_pos(h), list(cur)
    finally:
        R.refd_destroy(h)
