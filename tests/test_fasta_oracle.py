"""FASTA reader: the oracle's restatement of kseq_read_lossless_fasta (oracle/mbgc_oracle.c: mbo_fasta_parse) against
the reference's own utils/kseq.h instantiated over an in-memory file (oracle/_ref/libref_kseq.so) -- records, header
bytes, sequences, line-length state and the final status, for well-formed files and every early exit."""
import ctypes as C
import os

import numpy as np
import pytest

from fasta_cases import fasta_cases
from util import REF_DIR, oracle_lib, u8p, u64p

REFK = os.path.join(REF_DIR, "libref_kseq.so")


def oracle_parse(data):
    lib = oracle_lib()
    lib.mbo_fasta_parse.restype = C.c_uint64
    lib.mbo_fasta_parse.argtypes = [C.c_void_p, C.c_uint64, C.c_void_p, C.c_uint64, C.c_void_p, u64p, u64p, C.POINTER(C.c_int)]
    buf = np.frombuffer(data, dtype=np.uint8) if data else np.zeros(1, np.uint8)
    cap = data.count(b">") + 2
    recs = np.zeros((cap, 4), dtype=np.uint64)
    seq = np.zeros(max(1, len(data)), dtype=np.uint8)
    dna, mx, st = C.c_uint64(), C.c_uint64(), C.c_int()
    n = lib.mbo_fasta_parse(buf.ctypes.data, len(data), recs.ctypes.data, cap, seq.ctypes.data, C.byref(dna), C.byref(mx), C.byref(st))
    out = []
    for h_off, h_len, s_off, s_len in recs[:n]:
        out.append((data[int(h_off):int(h_off + h_len)], bytes(seq[int(s_off):int(s_off + s_len)])))
    return out, dna.value, mx.value, st.value


def kseq_parse(data):
    lib = C.CDLL(REFK)
    lib.refk_fasta_parse.restype = C.c_uint64
    lib.refk_fasta_parse.argtypes = [C.c_void_p, C.c_uint64, C.c_void_p, C.c_uint64, C.c_void_p, C.c_void_p, u64p, u64p, C.POINTER(C.c_int)]
    buf = np.frombuffer(data, dtype=np.uint8) if data else np.zeros(1, np.uint8)
    cap = data.count(b">") + 2
    recs = np.zeros((cap, 4), dtype=np.uint64)
    seq = np.zeros(max(1, len(data)), dtype=np.uint8)
    hdr = np.zeros(max(1, len(data)), dtype=np.uint8)
    dna, mx, st = C.c_uint64(), C.c_uint64(), C.c_int()
    n = lib.refk_fasta_parse(buf.ctypes.data, len(data), recs.ctypes.data, cap, seq.ctypes.data, hdr.ctypes.data, C.byref(dna), C.byref(mx), C.byref(st))
    out, h = [], 0
    for h_len, s_off, s_len, _ in recs[:n]:
        out.append((bytes(hdr[h:h + int(h_len)]), bytes(seq[int(s_off):int(s_off + s_len)])))
        h += int(h_len)
    return out, dna.value, mx.value, st.value


@pytest.mark.skipif(not os.path.exists(REFK), reason="oracle/_ref/libref_kseq.so not built (no /root/reference here)")
@pytest.mark.parametrize("name", sorted(fasta_cases()))
def test_oracle_reader_equals_kseq(name):
    data = fasta_cases()[name]
    got, want = oracle_parse(data), kseq_parse(data)
    assert got[3] == want[3], f"status {got[3]} vs kseq {want[3]}"
    assert len(got[0]) == len(want[0])
    for k, (a, b) in enumerate(zip(got[0], want[0])):
        assert a == b, f"record {k} differs"
    assert got[1:3] == want[1:3], "line-length state differs"


def test_statuses_cover_every_exit():
    seen = {oracle_parse(d)[3] for d in fasta_cases().values()}
    assert seen == {-1, -3, -4}
