"""Shared test helpers: ctypes bindings of the CPU oracle (oracle/libmbgc_oracle.so), of the reference
driver (oracle/_ref/libref_matcher.so, optional), FASTA reading and the seeded synthetic collections of
SURVEY.md 8d.  Test infrastructure only -- nothing here is imported by the product package."""
import ctypes as C
import os
import subprocess

import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
ORACLE_DIR = os.path.join(ROOT, "oracle")
REF_DIR = os.path.join(ORACLE_DIR, "_ref")
FIXTURES = os.path.join(REF_DIR, "fixtures")
GOLDEN = os.path.join(ROOT, "tests", "golden")
NO_LOCK = 2**64 - 1

u8p = C.POINTER(C.c_uint8)
u64p = C.POINTER(C.c_uint64)


def as_u8(buf):
    """numpy uint8 array (C-contiguous) from bytes/bytearray/ndarray."""
    if isinstance(buf, np.ndarray):
        return np.ascontiguousarray(buf, dtype=np.uint8)
    return np.frombuffer(bytes(buf), dtype=np.uint8)


def ptr(a):
    return a.ctypes.data_as(u8p)


# ----------------------------------------------------------------------------------------------
# oracle (our C restatement)
# ----------------------------------------------------------------------------------------------

class PackParams(C.Structure):
    _fields_ = [(n, C.c_int) for n in (
        "enableExtensionsWithMismatches", "gapBreakingMatchMinLength", "gapDepthOffsetEncoding",
        "gapDepthMismatchesEncoding", "lazyDecompressionSupport", "frugal64bitLenEncoding",
        "enable40bitReference", "mismatchesWithExclusion", "mmsMismatchesInitialScore",
        "mmsMismatchesScoreThreshold", "mmsMismatchPenalty", "mmsMatchBonus", "k", "unmatchedFractionFactor")]

    @staticmethod
    def for_mode(mode, enable40bit=False, lazy=True):
        # mbgccoder/MBGC_Params.h:66-97 defaults, :886-922 presets
        return PackParams(1, 256, 64, 2, 1 if lazy else 0, 0 if mode == 0 else 1, 1 if enable40bit else 0,
                          1, 125, 500, 50, 50, 32, 128)


def build_oracle():
    subprocess.run(["make", "-s", "oracle"], cwd=ORACLE_DIR, check=True)


_oracle = None


def oracle_lib():
    global _oracle
    if _oracle is None:
        path = os.path.join(ORACLE_DIR, "libmbgc_oracle.so")
        src = os.path.join(ORACLE_DIR, "mbgc_oracle.c")
        if not os.path.exists(path) or os.path.getmtime(path) < os.path.getmtime(src):
            build_oracle()
        L = C.CDLL(path)
        L.mbo_hash.restype = C.c_uint32
        L.mbo_hash.argtypes = [u8p, C.c_int]
        L.mbo_kmer_from_L.restype = C.c_int
        L.mbo_kmer_from_L.argtypes = [C.c_uint32, C.c_uint32]
        L.mbo_hash_size.restype = C.c_uint32
        L.mbo_hash_size.argtypes = [C.c_uint64, C.c_int]
        L.mbo_upper_reverse_complement.argtypes = [u8p, C.c_uint64, u8p]
        L.mbo_insert_ranges.restype = C.c_int
        L.mbo_insert_ranges.argtypes = [C.c_uint64, C.c_int64, C.c_int, C.c_int, C.POINTER(C.c_int64), C.c_int, u64p]
        L.mbo_create.restype = C.c_void_p
        L.mbo_create.argtypes = [C.c_uint64, C.c_uint32, C.c_int, C.c_int, C.c_int]
        L.mbo_destroy.argtypes = [C.c_void_p]
        for f in ("mbo_disable_sliding_window", "mbo_disable_circular_buffer"):
            getattr(L, f).argtypes = [C.c_void_p]
        L.mbo_set_sliding_window_size.argtypes = [C.c_void_p, C.c_int]
        L.mbo_load_ref.argtypes = [C.c_void_p, u8p, C.c_uint64, C.c_int, C.c_int, C.c_uint8]
        L.mbo_load_separator.argtypes = [C.c_void_p, C.c_uint8]
        L.mbo_match.restype = C.c_uint64
        L.mbo_match.argtypes = [C.c_void_p, u8p, C.c_uint64, C.c_uint32, C.c_uint64, C.c_int]
        L.mbo_matches.restype = C.c_void_p
        L.mbo_matches.argtypes = [C.c_void_p]
        L.mbo_acquire_lock.restype = C.c_uint64
        L.mbo_acquire_lock.argtypes = [C.c_void_p]
        L.mbo_release_lock.restype = C.c_int
        L.mbo_release_lock.argtypes = [C.c_void_p, C.c_uint64]
        L.mbo_ref.restype = C.c_void_p
        L.mbo_ref.argtypes = [C.c_void_p]
        for f in ("mbo_ref_length", "mbo_loading_position", "mbo_loaded_ref_length", "mbo_max_ref_length",
                  "mbo_table_size", "mbo_sampling_pos"):
            getattr(L, f).restype = C.c_uint64
            getattr(L, f).argtypes = [C.c_void_p]
        L.mbo_set_position.argtypes = [C.c_void_p, C.c_uint64, C.c_int]
        L.mbo_table.restype = C.c_void_p
        L.mbo_table.argtypes = [C.c_void_p]
        L.mbo_kmer.restype = C.c_int
        L.mbo_kmer.argtypes = [C.c_void_p]
        L.mbo_last_stats.argtypes = [C.c_void_p, u64p]
        L.mbo_streams_create.restype = C.c_void_p
        L.mbo_streams_destroy.argtypes = [C.c_void_p]
        L.mbo_stream_data.restype = C.c_void_p
        L.mbo_stream_data.argtypes = [C.c_void_p, C.c_int]
        L.mbo_stream_size.restype = C.c_uint64
        L.mbo_stream_size.argtypes = [C.c_void_p, C.c_int]
        L.mbo_streams_clear.argtypes = [C.c_void_p]
        L.mbo_after_sequence.argtypes = [C.c_void_p]
        L.mbo_after_target.argtypes = [C.c_void_p, C.POINTER(PackParams)]
        L.mbo_process_matches.restype = C.c_uint64
        L.mbo_process_matches.argtypes = [C.c_void_p, C.POINTER(PackParams), C.c_void_p, C.c_void_p, u64p, u8p,
                                          C.c_uint64, C.c_uint64, u64p, C.c_uint64]
        L.mbo_mismatch2code.restype = C.c_uint8
        L.mbo_mismatch2code.argtypes = [C.c_uint8, C.c_uint8]
        _oracle = L
    return _oracle


class _MatcherBase:
    """Common python face of the three matcher implementations (oracle, reference driver, CUDA)."""

    def match(self, text, min_len=32, lock=NO_LOCK):
        raise NotImplementedError

    def load_ref(self, text, rc=False, add_sep=True, sep=0):
        raise NotImplementedError


class OracleMatcher(_MatcherBase):
    def __init__(self, limit, L=32, k1=16, k2=1, skip_margin=16):
        self.lib = oracle_lib()
        self.h = self.lib.mbo_create(limit, L, k1, k2, skip_margin)

    def close(self):
        if self.h:
            self.lib.mbo_destroy(self.h)
            self.h = None

    def __del__(self):
        self.close()

    def disable_sliding_window(self):
        self.lib.mbo_disable_sliding_window(self.h)

    def disable_circular_buffer(self):
        self.lib.mbo_disable_circular_buffer(self.h)

    def set_sliding_window_size(self, f):
        self.lib.mbo_set_sliding_window_size(self.h, f)

    def load_ref(self, text, rc=False, add_sep=True, sep=0):
        a = as_u8(text)
        self.lib.mbo_load_ref(self.h, ptr(a), a.size, int(rc), int(add_sep), sep)

    def load_separator(self, sep=0):
        self.lib.mbo_load_separator(self.h, sep)

    def match(self, text, min_len=32, lock=NO_LOCK, prefilter=True):
        a = as_u8(text)
        n = self.lib.mbo_match(self.h, ptr(a), a.size, min_len, lock, int(prefilter))
        if n == NO_LOCK:
            raise ValueError("minMatchLength < K")
        if n == 0:
            return np.zeros((0, 3), dtype=np.uint64)
        p = self.lib.mbo_matches(self.h)
        return np.ctypeslib.as_array(C.cast(p, u64p), shape=(n, 3)).copy()

    def stats(self):
        out = (C.c_uint64 * 4)()
        self.lib.mbo_last_stats(self.h, out)
        return list(out)

    def acquire_lock(self):
        return self.lib.mbo_acquire_lock(self.h)

    def release_lock(self, v):
        return self.lib.mbo_release_lock(self.h, v)

    def set_position(self, pos, reached):
        self.lib.mbo_set_position(self.h, pos, reached)

    ref_length = property(lambda s: s.lib.mbo_ref_length(s.h))
    loading_position = property(lambda s: s.lib.mbo_loading_position(s.h))
    loaded_ref_length = property(lambda s: s.lib.mbo_loaded_ref_length(s.h))
    max_ref_length = property(lambda s: s.lib.mbo_max_ref_length(s.h))
    table_size = property(lambda s: s.lib.mbo_table_size(s.h))
    sampling_pos = property(lambda s: s.lib.mbo_sampling_pos(s.h))
    kmer = property(lambda s: s.lib.mbo_kmer(s.h))

    def table(self):
        return np.ctypeslib.as_array(C.cast(self.lib.mbo_table(self.h), C.POINTER(C.c_uint32)),
                                     shape=(self.table_size,))

    def ref(self, n=None):
        n = self.max_ref_length if n is None else n
        return np.ctypeslib.as_array(C.cast(self.lib.mbo_ref(self.h), u8p), shape=(n,))


# ----------------------------------------------------------------------------------------------
# reference driver (the reference's own class, built by oracle/Makefile when /root/reference exists)
# ----------------------------------------------------------------------------------------------

_refdrv = None


def have_ref_driver():
    return os.path.exists(os.path.join(REF_DIR, "libref_matcher.so"))


def have_ref_binaries():
    return os.path.exists(os.path.join(REF_DIR, "mbgc")) and os.path.exists(os.path.join(REF_DIR, "mbgc-dev"))


def have_fixtures():
    return os.path.exists(os.path.join(FIXTURES, "seqlist.txt"))


def refdrv_lib():
    global _refdrv
    if _refdrv is None:
        os.environ.setdefault("OMP_NUM_THREADS", "1")
        L = C.CDLL(os.path.join(REF_DIR, "libref_matcher.so"))
        L.refm_create.restype = C.c_void_p
        L.refm_create.argtypes = [C.c_uint64, C.c_uint32, C.c_int, C.c_int, C.c_int]
        L.refm_destroy.argtypes = [C.c_void_p]
        L.refm_disable_sliding_window.argtypes = [C.c_void_p]
        L.refm_disable_circular_buffer.argtypes = [C.c_void_p]
        L.refm_set_sliding_window_size.argtypes = [C.c_void_p, C.c_int]
        L.refm_load_ref.argtypes = [C.c_void_p, u8p, C.c_uint64, C.c_int, C.c_int, C.c_char]
        L.refm_load_separator.argtypes = [C.c_void_p, C.c_char]
        L.refm_match.restype = C.c_uint64
        L.refm_match.argtypes = [C.c_void_p, u8p, C.c_uint64, C.c_uint32, C.c_uint64]
        L.refm_get_matches.argtypes = [C.c_void_p, u64p]
        L.refm_ref.restype = C.c_void_p
        L.refm_ref.argtypes = [C.c_void_p]
        for f in ("refm_ref_length", "refm_loading_position", "refm_loaded_ref_length", "refm_max_ref_length",
                  "refm_acquire_lock", "refm_table_size", "refm_sampling_pos"):
            getattr(L, f).restype = C.c_uint64
            getattr(L, f).argtypes = [C.c_void_p]
        L.refm_release_lock.argtypes = [C.c_void_p, C.c_uint64]
        L.refm_set_position.argtypes = [C.c_void_p, C.c_uint64, C.c_int]
        L.refm_table.restype = C.c_void_p
        L.refm_table.argtypes = [C.c_void_p]
        L.refm_reached_count.restype = C.c_int
        L.refm_reached_count.argtypes = [C.c_void_p]
        L.refm_kmer.restype = C.c_int
        L.refm_kmer.argtypes = [C.c_void_p]
        _refdrv = L
    return _refdrv


class RefMatcher(_MatcherBase):
    """The reference's own SlidingWindow(Exp)SparseEMMatcher."""

    def __init__(self, limit, L=32, k1=16, k2=1, skip_margin=16):
        self.lib = refdrv_lib()
        self.h = self.lib.refm_create(limit, L, k1, k2, skip_margin)

    def close(self):
        if self.h:
            self.lib.refm_destroy(self.h)
            self.h = None

    def __del__(self):
        self.close()

    def disable_sliding_window(self):
        self.lib.refm_disable_sliding_window(self.h)

    def disable_circular_buffer(self):
        self.lib.refm_disable_circular_buffer(self.h)

    def set_sliding_window_size(self, f):
        self.lib.refm_set_sliding_window_size(self.h, f)

    def load_ref(self, text, rc=False, add_sep=True, sep=0):
        a = as_u8(text)
        self.lib.refm_load_ref(self.h, ptr(a), a.size, int(rc), int(add_sep), bytes([sep]))

    def load_separator(self, sep=0):
        self.lib.refm_load_separator(self.h, bytes([sep]))

    def match(self, text, min_len=32, lock=NO_LOCK):
        a = as_u8(text)
        n = self.lib.refm_match(self.h, ptr(a), a.size, min_len, lock)
        out = np.zeros((n, 3), dtype=np.uint64)
        if n:
            self.lib.refm_get_matches(self.h, out.ctypes.data_as(u64p))
        return out

    def acquire_lock(self):
        return self.lib.refm_acquire_lock(self.h)

    def release_lock(self, v):
        self.lib.refm_release_lock(self.h, v)

    def set_position(self, pos, reached):
        self.lib.refm_set_position(self.h, pos, reached)

    ref_length = property(lambda s: s.lib.refm_ref_length(s.h))
    loading_position = property(lambda s: s.lib.refm_loading_position(s.h))
    loaded_ref_length = property(lambda s: s.lib.refm_loaded_ref_length(s.h))
    max_ref_length = property(lambda s: s.lib.refm_max_ref_length(s.h))
    table_size = property(lambda s: s.lib.refm_table_size(s.h))
    sampling_pos = property(lambda s: s.lib.refm_sampling_pos(s.h))
    kmer = property(lambda s: s.lib.refm_kmer(s.h))

    def table(self):
        return np.ctypeslib.as_array(C.cast(self.lib.refm_table(self.h), C.POINTER(C.c_uint32)),
                                     shape=(self.table_size,))

    def ref(self, n=None):
        n = self.max_ref_length if n is None else n
        return np.ctypeslib.as_array(C.cast(self.lib.refm_ref(self.h), u8p), shape=(n,))


# ----------------------------------------------------------------------------------------------
# FASTA + synthetic collections
# ----------------------------------------------------------------------------------------------

def read_fasta(path):
    """[(header_without_'>', sequence_bytes_without_newlines)] -- what utils/kseq.h hands the matcher."""
    out = []
    with open(path, "rb") as f:
        data = f.read()
    for rec in data.split(b">")[1:]:
        nl = rec.find(b"\n")
        out.append((rec[:nl].decode(), rec[nl + 1:].replace(b"\n", b"").replace(b"\r", b"")))
    return out


def write_fasta(path, records, width=80):
    with open(path, "wb") as f:
        for header, seq in records:
            f.write(b">" + header.encode() + b"\n")
            seq = bytes(seq)
            for i in range(0, len(seq), width):
                f.write(seq[i:i + width] + b"\n")


from mbgc_b200.synth import synth_base, synth_mutant, synth_collection, mgmp_ref_limit  # noqa: E402,F401


# ----------------------------------------------------------------------------------------------
# row a12 (MBGC_Encoder::processLiteral): a seeded collection whose contigs are NOT proper for a whole-contig reference
# extension (few unmatched bases) but carry literals longer than the threshold -- novel islands shared between genomes
# ----------------------------------------------------------------------------------------------
LITERAL_EXT_CASES = ((1, 64, 8), (1, 200, 0), (0, 32, 16), (2, 100, 4))   # (mode, -N minimal literal length, -F context)


def literal_ext_genomes(n=10, length=200_000, seed=99):
    import numpy as np
    from mbgc_b200.synth import synth_base, synth_mutant
    rng = np.random.default_rng(seed + 7)
    base = synth_base(length, seed)
    acgt = np.frombuffer(b"ACGT", dtype=np.uint8)
    islands = [acgt[rng.integers(0, 4, size=int(rng.integers(100, 500)))] for _ in range(6)]
    out = [base.copy()]
    for i in range(1, n):
        g = synth_mutant(base, seed * 1000003 + i, 1e-4, 1e-5)
        for k in rng.choice(len(islands), size=2, replace=False):
            p = int(rng.integers(1000, g.size - 1000))
            g = np.concatenate([g[:p], islands[k], g[p:]])
        out.append(np.ascontiguousarray(g))
    return out


def literal_ext_collection(d, n=10, length=200_000, seed=99):
    """writes the collection as FASTA files + list.txt into d, returns the file names"""
    from mbgc_b200.synth import write_fasta
    names = []
    for i, g in enumerate(literal_ext_genomes(n, length, seed)):
        name = f"g{i:04d}.fna"
        write_fasta(os.path.join(d, name), f"syn_{i:04d} genome {i} with islands", g)
        names.append(name)
    with open(os.path.join(d, "list.txt"), "w") as f:
        f.write("\n".join(names) + "\n")
    return names
