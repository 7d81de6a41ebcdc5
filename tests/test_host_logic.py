"""Host-side logic and the C ABI surface, no GPU needed: the library loads, exports every symbol the header
declares, refuses to run without a device; index sampling ranges; reference-limit arithmetic."""
import ctypes as C
import os
import re

import numpy as np
import pytest

import mbgc_b200
from mbgc_b200.synth import fasta_bytes, mgmp_ref_limit, synth_collection, write_fasta
from util import ROOT, oracle_lib, read_fasta


def _cuda():
    try:
        import torch
        return torch.cuda.is_available()
    except Exception:
        return False


def test_library_exports_every_declared_symbol():
    lib = mbgc_b200.load_library()
    hdr = open(os.path.join(ROOT, "include", "mbgc_b200.h")).read()
    names = sorted(set(re.findall(r"\b(mbgcg_[a-z0-9_]+)\s*\(", hdr)))
    assert len(names) >= 25
    for n in names:
        assert hasattr(lib, n), f"{n} declared in include/mbgc_b200.h but not exported"


def test_oracle_is_not_linked_into_the_product():
    """the product library must not contain (or depend on) the CPU oracle"""
    import subprocess
    out = subprocess.run(["nm", "-D", "--defined-only", mbgc_b200.lib_path()], capture_output=True, text=True).stdout
    assert "mbo_" not in out
    deps = subprocess.run(["ldd", mbgc_b200.lib_path()], capture_output=True, text=True).stdout
    assert "oracle" not in deps
    for root, _, files in os.walk(os.path.join(ROOT, "mbgc_b200")):
        for f in files:
            if f.endswith((".py", ".cpp", ".cu", ".h", ".hpp")):
                src = open(os.path.join(root, f)).read()
                assert "mbgc_oracle" not in src and "libref_matcher" not in src, f"{f} references the oracle"


@pytest.mark.skipif(_cuda(), reason="only meaningful on a box without a GPU")
def test_fails_loudly_without_a_device():
    with pytest.raises(mbgc_b200.MatcherError) as e:
        mbgc_b200.Matcher(1 << 24)
    assert e.value.code == -1 and "no CPU fallback" in str(e.value)


def _ranges(sampling, pos1, K=28, k1=16):
    lib = oracle_lib()
    r = (C.c_int64 * 8)()
    new = C.c_uint64()
    n = lib.mbo_insert_ranges(sampling, pos1, K, k1, r, 4, C.byref(new))
    return [(r[2 * i], r[2 * i + 1]) for i in range(n)], new.value


def test_index_sampling_ranges_known_cases():
    """SURVEY.md 0.8 / Appendix B hand cases of matching/SlidingWindowSparseEMMatcher.cpp:146-171"""
    # refSamplingPos 4 999 984, refPos1 5 000 001 -> 5 100 001 leaves [5 098 288, 5 099 536) un-indexed
    rg, new = _ranges(4_999_984, 5_100_001)
    chunk_end = rg[0][1]
    tail_begin = rg[1][0]
    assert chunk_end == 5_098_288 and tail_begin == 5_099_536
    assert new == 5_099_984
    # a first call with refPos1 = 22 565 skips [20 496, 22 544)
    rg, new = _ranges(16, 22_565)
    assert rg == [(16, 20_496)] and new == 22_544  # the tail loop starts at 22 544 and finds nothing to sample yet
    # tiny reference: nothing to sample, sampling position moves to k1
    rg, new = _ranges(16, 20)
    assert rg == [] and new == 16
    # literal emulation of the two loops for random inputs
    rng = np.random.default_rng(3)
    for _ in range(300):
        k1 = int(rng.choice([4, 5, 7, 8, 16, 32]))
        K = int(rng.choice([12, 20, 28, 44]))
        pos1 = int(rng.integers(1, 300_000))
        sp = int(rng.integers(1, pos1 + 1))
        want, i1, kM = [], sp, k1 * 128
        while i1 < (pos1 - K) - kM:
            want += list(range(i1, i1 + kM, k1))
            i1 += kM
        t = k1 + int((pos1 - K - 1) / kM) * kM
        while t < pos1 - K + 1:
            want.append(t)
            t += k1
        rg, new = _ranges(sp, pos1, K, k1)
        got = [p for a, b in rg for p in range(a, b, k1)]
        assert got == want and new == t


def test_reference_limit_arithmetic():
    """MultipleGenomeMatchingProcessor.cpp:130-134,154-168: factor = 2^clamp(15 - clz32(n)/3, 5, 12)"""
    assert mgmp_ref_limit(3, 3_094_895) == 32 * 3_094_895 * 2      # Listeria: factor 32
    assert mgmp_ref_limit(100, 5_000_000) == 128 * 5_000_000 * 2   # 1.28 GB (C2)
    assert mgmp_ref_limit(1000, 4_900_000) == 256 * 4_900_000 * 2  # 2.5 GB (C3)
    assert mgmp_ref_limit(10, 1000) == 64 * (1 << 21) * 2          # MIN_BASIC_BLOCK_SIZE floor
    assert mgmp_ref_limit(10**6, 10**7) == 2**32 - 1               # capped below 4 GiB (40-bit mode is out of scope)


def test_synthetic_generator_and_fasta_io(tmp_path):
    gs = synth_collection(3, 50_000, 7, 1e-3, 1e-4)
    assert gs[0].size == 50_000 and set(np.unique(gs[0])) <= set(b"ACGT")
    assert np.array_equal(gs[1], synth_collection(3, 50_000, 7, 1e-3, 1e-4)[1])  # seeded
    p = tmp_path / "g.fna"
    write_fasta(str(p), "syn_0001 test", gs[1])
    (hdr, seq), = read_fasta(str(p))
    assert hdr == "syn_0001 test" and seq == gs[1].tobytes()
    assert os.path.getsize(p) == fasta_bytes(gs[1].size, len("syn_0001 test"))


def test_rc_lut_matches_reference_quirks():
    """upper-casing complement incl. IUPAC and the 0x7F -> 0 hole of the reference LUT (utils/helper.cpp:312-358)"""
    lib = oracle_lib()
    src = np.arange(256, dtype=np.uint8)
    dst = np.zeros(256, dtype=np.uint8)
    lib.mbo_upper_reverse_complement(src.ctypes.data_as(C.POINTER(C.c_uint8)), 256, dst.ctypes.data_as(C.POINTER(C.c_uint8)))
    lut = dst[::-1]
    assert bytes(lut[[ord(c) for c in "ACGTNacgtnUuYRKMBDHVWS"]]) == b"TGCANTGCANAARYMKVHDBSW"
    assert lut[0x7F] == 0 and lut[0] == 0 and lut[ord("X")] == ord("X") and lut[200] == 200


def test_vector_reverse_complement_equals_oracle():
    """the AVX2 reverse complement loadRef uses (mbgc_b200/csrc/hostpack.cpp) against the oracle's restatement of
    upperReverseComplement (utils/helper.cpp:405-410, table :312-358) -- every byte value, lengths around the 32-byte
    blocks, random texts with lower case / IUPAC / N / 0x7F"""
    import ctypes as C
    from util import oracle_lib
    lib = mbgc_b200.load_library()
    lib.mbgcg_selftest_revcomp.argtypes = [C.c_void_p, C.c_uint64, C.c_void_p]
    o = oracle_lib()
    rng = np.random.default_rng(8)
    alphabet = np.frombuffer(b"ACGTNacgtnUuYyRrKkMmBbDdHhVvWwSsXx*-\x7f\x00\xff", dtype=np.uint8)
    cases = [np.arange(256, dtype=np.uint8), np.tile(np.arange(256, dtype=np.uint8), 3)[5:]]
    for n in (0, 1, 31, 32, 33, 63, 64, 65, 95, 96, 97, 1000, 100_003):
        cases.append(alphabet[rng.integers(0, alphabet.size, n)].copy())
        cases.append(rng.integers(0, 256, n).astype(np.uint8))
    for src in cases:
        want = np.zeros(max(1, src.size), dtype=np.uint8)
        got = np.full(max(1, src.size), 0xEE, dtype=np.uint8)
        if src.size:
            o.mbo_upper_reverse_complement(src.ctypes.data_as(C.POINTER(C.c_uint8)), src.size, want.ctypes.data_as(C.POINTER(C.c_uint8)))
        lib.mbgcg_selftest_revcomp(src.ctypes.data, src.size, got.ctypes.data)
        assert np.array_equal(got[:src.size], want[:src.size]), f"length {src.size}"


def test_host_packer_round_trip():
    """2-bit packing of the text transfer path (mbgc_b200/csrc/hostpack.cpp): exact ASCII comes back, including
    N runs, IUPAC codes, lower case and lengths that are not multiples of 4 or 32; texts that are mostly
    non-ACGT are refused (they travel raw)."""
    import ctypes as C
    lib = mbgc_b200.load_library()
    rng = np.random.default_rng(5)
    acgt = np.frombuffer(b"ACGT", dtype=np.uint8)
    for n in (0, 1, 3, 4, 5, 31, 32, 33, 63, 64, 65, 1000, 262_144, 262_147):
        t = acgt[rng.integers(0, 4, n)].copy()
        if n > 40:
            t[7:19] = ord("N")
            t[n // 2] = ord("R")
            t[n - 1] = ord("n")
            t[n - 3] = 0
        if n > 100_000:
            t[50_000:58_000] = ord("N")                 # long run: one exception entry
            t[90_000:90_040] = np.frombuffer(b"acgtn" * 8, dtype=np.uint8)
        out = np.zeros(n, dtype=np.uint8)
        rc = lib.mbgcg_selftest_pack(t.ctypes.data, n, out.ctypes.data)
        assert rc == 0
        assert np.array_equal(out, t), f"length {n}"
    lower = np.frombuffer(bytes(acgt[rng.integers(0, 4, 10_000)]).lower(), dtype=np.uint8).copy()
    out = np.zeros(lower.size, dtype=np.uint8)
    assert lib.mbgcg_selftest_pack(lower.ctypes.data, lower.size, out.ctypes.data) == 1


def test_fasta_image_packer_round_trip():
    """mbgcg_image_pack (hostpack.cpp: pack_fasta_image): whatever the file holds, header + codes + exception runs rebuild it
    byte for byte; the line ends of a well-formed file leave as one run per 512 lines, so a genome travels at a quarter of its
    size; files with too many exceptional bytes are refused (they travel as they are)"""
    import ctypes as C
    lib = mbgc_b200.load_library()
    lib.mbgcg_selftest_pack_image.restype = C.c_int
    lib.mbgcg_selftest_pack_image.argtypes = [C.c_void_p, C.c_uint64, C.c_void_p, C.POINTER(C.c_uint64), C.POINTER(C.c_uint64)]
    rng = np.random.default_rng(3)
    acgt = np.frombuffer(b"ACGT", dtype=np.uint8)

    def fasta(nrec, length, width, final=True, eol=b"\n"):
        parts = []
        for r in range(nrec):
            parts.append(f">rec{r} some header text | x=1".encode() + eol)
            s = acgt[rng.integers(0, 4, length)].copy()
            if length > 2000:
                s[500:1500] = ord("N")
                s[1700] = ord("n")
                s[1800:1810] = np.frombuffer(b"acgtRYKMac", dtype=np.uint8)
            b = bytes(s)
            parts.append(eol.join(b[i:i + width] for i in range(0, len(b), width)))
            if final or r < nrec - 1:
                parts.append(eol)
        return np.frombuffer(b"".join(parts), dtype=np.uint8)

    def run(img):
        out = np.zeros(max(img.size, 1), dtype=np.uint8)
        pb, nr = C.c_uint64(), C.c_uint64()
        rc = lib.mbgcg_selftest_pack_image(img.ctypes.data, img.size, out.ctypes.data, C.byref(pb), C.byref(nr))
        return rc, out[:img.size], pb.value, nr.value

    big = fasta(1, 2_000_000, 80)
    rc, out, pb, nr = run(big)
    assert rc == 0 and np.array_equal(out, big)
    assert pb < 0.26 * big.size and nr < 200, (pb, nr)
    for img in (fasta(3, 100_003, 60), fasta(2, 7000, 80, final=False), fasta(1, 300_000, 1000),
                np.frombuffer(b">x\nACGT", dtype=np.uint8), np.frombuffer(b"\n\n\n\n>a\n\nAC\n\n\nGT\n", dtype=np.uint8),
                np.frombuffer(b">only a header", dtype=np.uint8)):
        rc, out, pb, nr = run(img)
        assert rc == 0 and np.array_equal(out, img), img.size
    assert run(fasta(2, 9000, 70, eol=b"\r\n"))[0] == 1      # two exceptional bytes per line: refused
    assert run(np.frombuffer(bytes(fasta(1, 20_000, 80)).lower(), dtype=np.uint8))[0] == 1


def test_header_is_plain_c(tmp_path):
    """include/mbgc_b200.h is the drop-in boundary for any host language: it must compile as C99 on its own"""
    import os
    import shutil
    import subprocess
    gcc = shutil.which("gcc") or "/usr/bin/gcc"
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    src = tmp_path / "hdr.c"
    src.write_text('#include "mbgc_b200.h"\nint main(void) { mbgcg_pack_params_t p; mbgcg_pack_result_t r; mbgcg_fasta_rec_t f; '
                   'mbgcg_match_t m; mbgcg_state_t s; (void) p; (void) r; (void) f; (void) m; (void) s; return MBGCG_OK; }\n')
    res = subprocess.run([gcc, "-std=c99", "-Wall", "-Wextra", "-pedantic", "-Werror", "-I", os.path.join(root, "include"),
                          "-c", str(src), "-o", str(tmp_path / "hdr.o")], capture_output=True, text=True)
    assert res.returncode == 0, res.stderr
