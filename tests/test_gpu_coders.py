"""Row a14 on the device (mbgc_b200/csrc/coders.cu through the C ABI: mbgcg_symbols_pack / _unpack, mbgcg_varlen_encode /
_decode) against the oracle restatement, which tests/test_coders_oracle.py pins against the reference's own classes."""
import numpy as np
import pytest

import mbgc_b200
from coders_cases import (ACGNT, ACGT, OWN_BOOK, SYMBOL_CASES, VARLEN_SIZES, o_decode, o_encode, o_pack, o_unpack,
                          oracle_coders, random_symbols, ref_book, ref_coders)

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def g():
    m = mbgc_b200.Matcher(1 << 22, 32, 16, 1, 16, device=0)
    yield m
    m.close()


@pytest.mark.parametrize("symbols,n,seed", SYMBOL_CASES + [(ACGT, 10_000_003, 5), (ACGNT, 9_000_001, 6)])
def test_symbols_pack_unpack(g, symbols, n, seed):
    L = oracle_coders()
    src = random_symbols(symbols, n, seed)
    want = o_pack(L, symbols, src)
    got = g.packSymbols(symbols, src)
    assert np.array_equal(got, want)
    for pos, length in ((0, n), (1, max(0, n - 1)), (n // 3, n // 2)):
        if pos + length > n or length == 0:
            continue
        back = g.unpackSymbols(symbols, got, pos, length)
        assert np.array_equal(back, src[pos:pos + length])
        if n <= 70000:
            assert np.array_equal(back, o_unpack(L, symbols, want, pos, length))


def test_symbols_reject_foreign_byte(g):
    with pytest.raises(mbgc_b200.MatcherError):
        g.packSymbols(ACGT, np.frombuffer(b"ACGTNACGT", dtype=np.uint8))


@pytest.mark.parametrize("n", (1, 2, 3) + VARLEN_SIZES + (5_000_017,))
def test_varlen_own_book(g, n):
    L = oracle_coders()
    src = random_symbols(b"ACGTN", n, 7 * n)
    want = o_encode(L, OWN_BOOK, src)
    got = g.varLenEncode(OWN_BOOK, src)
    assert np.array_equal(got, want)
    assert np.array_equal(g.varLenDecode(OWN_BOOK, got, n), src)
    with pytest.raises(mbgc_b200.MatcherError):
        g.varLenDecode(OWN_BOOK, got, n + 1)      # "Unexpected decoded length"


@pytest.mark.skipif(ref_coders() is None, reason="oracle/_ref/libref_coders.so not built (the static books are the reference's data)")
@pytest.mark.parametrize("book_id", [0, 1, 2])
def test_varlen_static_books_of_the_reference(g, book_id):
    L = oracle_coders()
    book = ref_book(ref_coders(), book_id)
    for n in (64, 1000, 16385, 3_000_001):
        src = random_symbols(b"ACGT", n, 100 * book_id + n, run_of=ord("N"))
        want = o_encode(L, book, src)
        got = g.varLenEncode(book, src)
        assert np.array_equal(got, want)
        assert np.array_equal(g.varLenDecode(book, got, n), src)


def test_varlen_errors(g):
    with pytest.raises(mbgc_b200.MatcherError):
        g.varLenEncode(OWN_BOOK, np.frombuffer(b"ACGTXACGT" * 20, dtype=np.uint8))   # unexpected symbol on the parse path
    with pytest.raises(mbgc_b200.MatcherError):
        g.varLenEncode("AC\nA", np.frombuffer(b"AAAA", dtype=np.uint8))               # first word must be a single symbol
