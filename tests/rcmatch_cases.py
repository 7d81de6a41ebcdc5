"""Inputs shared by the CPU and GPU tests of row f3 (SimpleSequenceMatcher::rcMatchSequence over CopMEMMatcher).
Test infrastructure."""
import ctypes as C
import os

import numpy as np

from util import REF_DIR, oracle_lib

_COMP = np.arange(256, dtype=np.uint8)
for a, b in zip(b"ACGTNYRKMBDHVWS", b"TGCANRYMKVHDBSW"):
    _COMP[a] = b
    _COMP[a + 32] = b + 32
_COMP[0x7F] = 0


def revcomp(a):
    return _COMP[a[::-1]]


def rc_sequence(n, seed, n_repeats=40, max_len=4000, marks=True, mutate=True):
    """random ACGT with reverse-complement copies of earlier stretches pasted in (some with a few substitutions), a
    palindromic stretch, a long N run, and -- like the literal stream mode 3 hands over -- scattered marker / code bytes"""
    rng = np.random.default_rng(seed)
    s = np.frombuffer(b"ACGT", dtype=np.uint8)[rng.integers(0, 4, n)].copy()
    for _ in range(n_repeats if n > 2000 else 0):
        ln = int(rng.integers(30, min(max_len, n // 4)))
        src = int(rng.integers(0, n - ln))
        dst = int(rng.integers(0, n - ln))
        piece = revcomp(s[src:src + ln])
        if mutate and rng.random() < 0.5:
            for p in rng.integers(0, ln, max(1, ln // 300)):
                piece[p] = b"ACGT"[int(rng.integers(0, 4))]
        s[dst:dst + ln] = piece
    if n > 5000:
        h = int(rng.integers(60, 400))
        at = int(rng.integers(0, n - 2 * h))
        s[at + h:at + 2 * h] = revcomp(s[at:at + h])          # a reverse-complement palindrome: overlapping src / dst
        at = int(rng.integers(0, n - 700))
        s[at:at + 600] = ord("N")
    if marks and n > 100:
        for p in rng.integers(0, n, n // 500 + 1):
            s[p] = [0xA5, 0xA2, 0, 1, 2, 3, ord("n"), ord("a")][int(rng.integers(0, 8))]
    return s


RC_CASES = [(0, 1), (54, 2), (55, 3), (200, 4), (3000, 5), (20_000, 6), (300_000, 7), (1_200_000, 8)]


def ref_copmem():
    path = os.path.join(REF_DIR, "libref_copmem.so")
    if not os.path.exists(path):
        return None
    L = C.CDLL(path)
    L.refr_rc_match_sequence.restype = C.c_uint64
    L.refr_rc_match_sequence.argtypes = [C.c_void_p, C.c_uint64, C.c_uint64, C.c_void_p, C.POINTER(C.c_uint64), C.c_void_p,
                                         C.POINTER(C.c_uint64), C.c_int]
    L.refr_rc_restore.restype = C.c_uint64
    L.refr_rc_restore.argtypes = [C.c_void_p, C.c_uint64, C.c_void_p, C.c_uint64, C.c_void_p, C.c_uint64, C.c_uint64, C.c_void_p]
    return L


def oracle_rc():
    L = oracle_lib()
    L.mbo_rc_match_sequence.restype = C.c_uint64
    L.mbo_rc_match_sequence.argtypes = [C.c_void_p, C.c_uint64, C.c_uint32, C.c_uint32, C.c_void_p, C.POINTER(C.c_uint64),
                                        C.c_void_p, C.POINTER(C.c_uint64), C.POINTER(C.c_uint64)]
    return L


def o_rc_match(L, seq, target_len=55, min_len=2**32 - 1):
    s = seq.copy()
    off = np.zeros(4 * (s.size // 24 + 4), dtype=np.uint8)
    ln = np.zeros(10 * (s.size // 24 + 4), dtype=np.uint8)
    no, nl, raw = C.c_uint64(), C.c_uint64(), C.c_uint64()
    n = L.mbo_rc_match_sequence(s.ctypes.data, s.size, target_len, min_len, off.ctypes.data, C.byref(no), ln.ctypes.data,
                                C.byref(nl), C.byref(raw))
    if n == 2**64 - 1:
        return None
    return s[:n].copy(), off[:no.value].copy(), ln[:nl.value].copy(), int(raw.value)


def r_rc_match(L, seq, target_len=55, threads=1):
    s = seq.copy()
    off = np.zeros(4 * (s.size // 24 + 4), dtype=np.uint8)
    ln = np.zeros(10 * (s.size // 24 + 4), dtype=np.uint8)
    no, nl = C.c_uint64(), C.c_uint64()
    n = L.refr_rc_match_sequence(s.ctypes.data, s.size, target_len, off.ctypes.data, C.byref(no), ln.ctypes.data, C.byref(nl), threads)
    return s[:n].copy(), off[:no.value].copy(), ln[:nl.value].copy()


def r_rc_restore(L, s, off, ln, org_len):
    out = np.zeros(org_len + 64, dtype=np.uint8)
    n = L.refr_rc_restore(s.ctypes.data if s.size else None, s.size, off.ctypes.data if off.size else None, off.size,
                          ln.ctypes.data if ln.size else None, ln.size, org_len, out.ctypes.data)
    return out[:n].copy()
