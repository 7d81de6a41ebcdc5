"""Row a14 (SymbolsPackingFacility, VarLenDNACoder): the oracle restatement (oracle/mbgc_oracle.c: mbo_symbols_*,
mbo_varlen_*) against the REFERENCE's own classes compiled from where they lie (oracle/_ref/libref_coders.so) and against
golden digests generated from them (tests/golden/make_golden.py -> golden.json["coders"]).  CPU only."""
import hashlib
import json
import os

import numpy as np
import pytest

from coders_cases import (ACGNT, ACGT, OWN_BOOK, SYMBOL_CASES, VARLEN_SIZES, o_decode, o_encode, o_pack, o_unpack,
                          oracle_coders, random_symbols, ref_book, ref_coders)
from util import GOLDEN

REF = ref_coders()
needs_ref = pytest.mark.skipif(REF is None, reason="oracle/_ref/libref_coders.so not built (no /root/reference here)")


def sha(a):
    return hashlib.sha1(bytes(a)).hexdigest()[:16]


@needs_ref
@pytest.mark.parametrize("symbols,n,seed", SYMBOL_CASES)
def test_symbols_pack_and_unpack_equal_reference(symbols, n, seed):
    L = oracle_coders()
    src = random_symbols(symbols, n, seed)
    got = o_pack(L, symbols, src)
    want = np.zeros(n + 8, dtype=np.uint8)
    k = REF.refc_symbols_pack(symbols, len(symbols), src.ctypes.data, n, want.ctypes.data)
    assert np.array_equal(got, want[:k])
    for pos, length in ((0, n), (1, max(0, n - 1)), (n // 3, n // 2), (max(0, n - 2), min(2, n))):
        if pos + length > n:
            continue
        w = np.zeros(max(length, 1), dtype=np.uint8)
        REF.refc_symbols_unpack(symbols, len(symbols), got.ctypes.data, pos, length, w.ctypes.data)
        assert np.array_equal(o_unpack(L, symbols, got, pos, length), w[:length])
        assert np.array_equal(w[:length], src[pos:pos + length])


@needs_ref
@pytest.mark.parametrize("book_id", [0, 1, 2])
@pytest.mark.parametrize("n", VARLEN_SIZES)
def test_varlen_equals_reference(book_id, n):
    L = oracle_coders()
    book = ref_book(REF, book_id)
    # the static books know A, C, G, T, N and two markers ('%', '"'); book 1 has no word for a lone marker behind N etc.
    src = random_symbols(b"ACGT", n, 100 * book_id + n, run_of=ord("N"))
    got = o_encode(L, book, src)
    want = np.zeros(n + 8, dtype=np.uint8)
    k = REF.refc_varlen_encode(book.encode("latin-1"), src.ctypes.data, n, want.ctypes.data)
    assert got is not None and np.array_equal(got, want[:k])
    back = np.zeros(n, dtype=np.uint8)
    REF.refc_varlen_decode(book.encode("latin-1"), got.ctypes.data, got.size, back.ctypes.data, n)
    assert np.array_equal(back, src)
    m, dec = o_decode(L, book, got, n)
    assert m == n and np.array_equal(dec, src)


def test_varlen_own_book_round_trip_and_tail_rules():
    L = oracle_coders()
    for n in (1, 2, 3) + VARLEN_SIZES:
        src = random_symbols(b"ACGTN", n, 7 * n)
        codes = o_encode(L, OWN_BOOK, src)
        assert codes is not None and len(codes) <= n
        m, dec = o_decode(L, OWN_BOOK, codes, n)
        assert m == n and np.array_equal(dec, src)
    assert o_encode(L, OWN_BOOK, np.frombuffer(b"ACGX", dtype=np.uint8)) is None      # unexpected symbol
    assert o_encode(L, "AC\nA", np.frombuffer(b"AAAA", dtype=np.uint8)) is None        # first word must be one symbol


def test_symbols_own_cases():
    L = oracle_coders()
    assert bytes(o_pack(L, ACGT, np.frombuffer(b"ACGTTT", dtype=np.uint8))) == bytes([0 * 64 + 1 * 16 + 2 * 4 + 3, 3 * 64 + 3 * 16])
    assert bytes(o_pack(L, ACGNT, np.frombuffer(b"NTA", dtype=np.uint8))) == bytes([3 * 25 + 4 * 5 + 0])
    assert o_pack(L, ACGT, np.frombuffer(b"ACGN", dtype=np.uint8)) is None


def test_golden_digests():
    """digests produced from the reference's classes by tests/golden/make_golden.py (so the pin travels without _ref)"""
    g = json.load(open(os.path.join(GOLDEN, "golden.json"))).get("coders")
    if g is None:
        pytest.skip("golden.json has no coders section yet")
    L = oracle_coders()
    for key, want in g["symbols"].items():
        alpha, n, seed = key.split(":")
        src = random_symbols(alpha.encode(), int(n), int(seed))
        assert sha(o_pack(L, alpha.encode(), src)) == want
    for key, want in g["varlen_own_book"].items():
        n = int(key)
        src = random_symbols(b"ACGTN", n, 7 * n)
        assert sha(o_encode(L, OWN_BOOK, src)) == want
