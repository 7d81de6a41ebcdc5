"""Inputs shared by the CPU and GPU tests of row a14 (SymbolsPackingFacility, VarLenDNACoder).  Test infrastructure."""
import ctypes as C
import os

import numpy as np

from util import REF_DIR, oracle_lib, u8p

ACGT = b"ACGT"
ACGNT = b"ACGNT"   # the reference's acgtnSymbols order (coders/SymbolsPackingFacility.cpp:13)

# a small book of our own: the first word a single symbol, words of 1..4 symbols, with aliases on the low 6 bits left out
OWN_BOOK = "\n".join(["A", "C", "G", "T", "N", "AA", "AC", "ACG", "ACGT", "TTTT", "GAT", "TA", "CCC", "GGGG", "NN", "AN", "NAC"])


def ref_coders():
    path = os.path.join(REF_DIR, "libref_coders.so")
    if not os.path.exists(path):
        return None
    L = C.CDLL(path)
    L.refc_symbols_pack.restype = C.c_uint64
    L.refc_symbols_pack.argtypes = [C.c_char_p, C.c_int, C.c_void_p, C.c_uint64, C.c_void_p]
    L.refc_symbols_unpack.argtypes = [C.c_char_p, C.c_int, C.c_void_p, C.c_uint64, C.c_uint64, C.c_void_p]
    L.refc_varlen_book.restype = C.c_uint64
    L.refc_varlen_book.argtypes = [C.c_int, C.c_void_p]
    L.refc_varlen_encode.restype = C.c_uint64
    L.refc_varlen_encode.argtypes = [C.c_char_p, C.c_void_p, C.c_uint64, C.c_void_p]
    L.refc_varlen_decode.argtypes = [C.c_char_p, C.c_void_p, C.c_uint64, C.c_void_p, C.c_uint64]
    return L


def ref_book(L, book_id):
    buf = C.create_string_buffer(4096)
    n = L.refc_varlen_book(book_id, buf)
    return buf.raw[:n].decode("latin-1")


def oracle_coders():
    L = oracle_lib()
    L.mbo_symbols_pack.restype = C.c_uint64
    L.mbo_symbols_pack.argtypes = [C.c_char_p, C.c_int, C.c_void_p, C.c_uint64, C.c_void_p]
    L.mbo_symbols_unpack.argtypes = [C.c_char_p, C.c_int, C.c_void_p, C.c_uint64, C.c_uint64, C.c_void_p]
    L.mbo_varlen_encode.restype = C.c_uint64
    L.mbo_varlen_encode.argtypes = [C.c_char_p, C.c_void_p, C.c_uint64, C.c_void_p]
    L.mbo_varlen_decode.restype = C.c_uint64
    L.mbo_varlen_decode.argtypes = [C.c_char_p, C.c_void_p, C.c_uint64, C.c_void_p, C.c_uint64]
    return L


def random_symbols(alphabet, n, seed, run_of=None):
    rng = np.random.default_rng(seed)
    a = np.frombuffer(alphabet, dtype=np.uint8)[rng.integers(0, len(alphabet), n)].copy()
    if run_of is not None and n > 64:   # a few long runs of one symbol (N stretches)
        for s in rng.integers(0, n - 40, 5):
            a[s:s + int(rng.integers(1, 40))] = run_of
    return a


def o_pack(L, symbols, src):
    dst = np.zeros(len(src) + 8, dtype=np.uint8)
    n = L.mbo_symbols_pack(symbols, len(symbols), src.ctypes.data, src.size, dst.ctypes.data)
    return None if n == 2**64 - 1 else dst[:n].copy()


def o_unpack(L, symbols, packed, pos, length):
    dst = np.zeros(max(length, 1), dtype=np.uint8)
    L.mbo_symbols_unpack(symbols, len(symbols), packed.ctypes.data, pos, length, dst.ctypes.data)
    return dst[:length].copy()


def o_encode(L, book, src):
    dst = np.zeros(len(src) + 8, dtype=np.uint8)
    n = L.mbo_varlen_encode(book.encode("latin-1"), src.ctypes.data, src.size, dst.ctypes.data)
    return None if n == 2**64 - 1 else dst[:n].copy()


def o_decode(L, book, codes, cap):
    dst = np.zeros(max(cap, 1), dtype=np.uint8)
    n = L.mbo_varlen_decode(book.encode("latin-1"), codes.ctypes.data, codes.size, dst.ctypes.data, cap)
    return int(n), dst[:min(cap, n)].copy()


SYMBOL_CASES = [(ACGT, n, 11 + n) for n in (0, 1, 3, 4, 5, 1000, 65537)] + [(ACGNT, n, 23 + n) for n in (0, 2, 3, 4, 999, 65536)]
VARLEN_SIZES = (4, 5, 7, 63, 64, 65, 1000, 16384, 16385, 200_003)
