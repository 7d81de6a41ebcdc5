"""Regenerates tests/golden/golden.json from the REFERENCE itself (run in the container that has
/root/reference, after `make -C oracle ref fixtures`):

    python tests/golden/make_golden.py

Everything recorded here comes out of the reference's own code: oracle/_ref/libref_matcher.so (the
reference's SlidingWindowSparseEMMatcher.cpp behind a C driver) and oracle/_ref/mbgc, mbgc-dev (the stock
CLI).  Inputs are the seeded scenarios of tests/scenarios.py and seeded synthetic FASTA collections, so the
tests can rebuild them anywhere; only digests are stored."""
import hashlib
import json
import os
import subprocess
import sys
import tempfile

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))

from scenarios import all_scenarios, replay  # noqa: E402
from util import FIXTURES, REF_DIR, RefMatcher, refdrv_lib  # noqa: E402
from mbgc_b200.synth import synth_collection, write_fasta  # noqa: E402


def digest(arrs):
    h = hashlib.sha256()
    for a in arrs:
        h.update(np.ascontiguousarray(a, dtype=np.uint64).tobytes())
        h.update(b"|")
    return h.hexdigest()


def golden_collection(d):
    """the small synthetic collection of the stream-level golden test"""
    gs = synth_collection(6, 200_000, 99, 5e-3, 5e-4)
    names = []
    for i, g in enumerate(gs):
        name = f"g{i:04d}.fna"
        write_fasta(os.path.join(d, name), f"syn_{i:04d} synthetic E. coli-like genome {i}", g)
        names.append(name)
    with open(os.path.join(d, "list.txt"), "w") as f:
        f.write("\n".join(names) + "\n")
    return names


def stream_digests(d, list_name, mode, extra=()):
    arch = os.path.join(d, f"m{mode}.mbgc")
    env = dict(os.environ)
    env.pop("OMP_NUM_THREADS", None)
    subprocess.run([os.path.join(REF_DIR, "mbgc-dev"), "c", "-t", "1", "-m", str(mode), *extra, "-f", list_name, arch],
                   cwd=d, check=True, capture_output=True, env=env)
    subprocess.run([os.path.join(REF_DIR, "mbgc-dev"), "v", "-Q", "-D", arch], cwd=d, check=True, capture_output=True,
                   env=env)
    dumps = sorted(f for f in os.listdir(d) if f.startswith(f"m{mode}.mbgc_dump_"))
    # the last six dumps are literals, locksPos, gapDelta, gapMismatchFlags, mapOff, mapLen
    names = ["literals", "locksPos", "gapDelta", "gapFlags", "mapOff", "mapLen"]
    out = {}
    for n, f in zip(names, dumps[-6:]):
        data = open(os.path.join(d, f), "rb").read()
        out[n] = {"size": len(data), "sha256": hashlib.sha256(data).hexdigest()}
    out["archive_md5"] = hashlib.md5(open(arch, "rb").read()).hexdigest()
    for f in dumps:
        os.remove(os.path.join(d, f))
    return out


def coders_golden():
    """row a14: digests of what the reference's own SymbolsPackingFacility / VarLenDNACoder produce
    (oracle/_ref/libref_coders.so) for the seeded inputs of tests/coders_cases.py"""
    from coders_cases import OWN_BOOK, SYMBOL_CASES, VARLEN_SIZES, random_symbols, ref_coders
    L = ref_coders()
    sha = lambda a: hashlib.sha1(bytes(a)).hexdigest()[:16]
    out = {"symbols": {}, "varlen_own_book": {}}
    for symbols, n, seed in SYMBOL_CASES:
        src = random_symbols(symbols, n, seed)
        dst = np.zeros(n + 8, dtype=np.uint8)
        k = L.refc_symbols_pack(symbols, len(symbols), src.ctypes.data, n, dst.ctypes.data)
        out["symbols"][f"{symbols.decode()}:{n}:{seed}"] = sha(dst[:k])
    for n in VARLEN_SIZES:
        src = random_symbols(b"ACGTN", n, 7 * n)
        dst = np.zeros(n + 8, dtype=np.uint8)
        k = L.refc_varlen_encode(OWN_BOOK.encode(), src.ctypes.data, n, dst.ctypes.data)
        out["varlen_own_book"][str(n)] = sha(dst[:k])
    return out


def rcmatch_golden():
    """row f3: digests of what the reference's own SimpleSequenceMatcher::rcMatchSequence produces
    (oracle/_ref/libref_copmem.so, single-threaded index build) for the seeded inputs of tests/rcmatch_cases.py"""
    from rcmatch_cases import RC_CASES, r_rc_match, rc_sequence, ref_copmem
    L = ref_copmem()
    out = {}
    for n, seed in RC_CASES:
        for tl in (55, 32):
            s, off, ln = r_rc_match(L, rc_sequence(n, seed), tl)
            h = hashlib.sha1()
            for a in (s, off, ln):
                h.update(bytes(a))
                h.update(b"|")
            out[f"{n}:{seed}:{tl}"] = h.hexdigest()[:16]
    return out


def literal_ext_golden():
    """row a12: stream digests of the developer build of the stock CLI with MBGC_Encoder::processLiteral switched on
    (`-N <minimal literal length> -F <context>`, main.cpp:308-313) on the seeded collection of tests/util.py"""
    from util import LITERAL_EXT_CASES, literal_ext_collection
    out = {}
    for mode, min_len, ctx in LITERAL_EXT_CASES:
        with tempfile.TemporaryDirectory() as d:
            literal_ext_collection(d)
            out[f"{mode}:{min_len}:{ctx}"] = stream_digests(d, "list.txt", mode, ("-N", str(min_len), "-F", str(ctx)))
            off = stream_digests(d, "list.txt", mode)
            assert off["literals"]["sha256"] != out[f"{mode}:{min_len}:{ctx}"]["literals"]["sha256"], "the option changed nothing"
    return out


def main():
    if "--only-literal-ext" in sys.argv:
        path = os.path.join(HERE, "golden.json")
        g = json.load(open(path))
        g["literal_ext"] = literal_ext_golden()
        with open(path, "w") as f:
            json.dump(g, f, indent=1, sort_keys=True)
        print("updated literal_ext section of", path)
        return
    if "--only-rcmatch" in sys.argv:
        path = os.path.join(HERE, "golden.json")
        g = json.load(open(path))
        g["rcmatch"] = rcmatch_golden()
        with open(path, "w") as f:
            json.dump(g, f, indent=1, sort_keys=True)
        print("updated rcmatch section of", path)
        return
    if "--only-coders" in sys.argv:   # refresh one section, keep the rest
        path = os.path.join(HERE, "golden.json")
        g = json.load(open(path))
        g["coders"] = coders_golden()
        with open(path, "w") as f:
            json.dump(g, f, indent=1, sort_keys=True)
        print("updated coders section of", path)
        return
    g = {"scenarios": {}, "streams": {}, "hash": [], "listeria": {}}
    # 1. match lists / table / cursor state of every scenario, from the reference class
    for sm in (16, 24):
        for sc in all_scenarios():
            r = RefMatcher(sc["limit"], 32, 16, 1, sm)
            res, state = replay(r, sc)
            g["scenarios"][f"{sc['name']}/{sm}"] = {
                "matches": digest(res), "counts": [int(len(x)) for x in res], "state": [int(x) for x in state],
                "table": hashlib.sha256(r.table().tobytes()).hexdigest()}
            r.close()
    # 2. stream-level known answers of the stock CLI on a seeded collection, modes 0..2 (mode 3 rewrites the literal
    #    stream with CopMEM afterwards, outside the path)
    with tempfile.TemporaryDirectory() as d:
        golden_collection(d)
        for mode in (0, 1, 2):
            g["streams"][str(mode)] = stream_digests(d, "list.txt", mode)
    # 3. the shipped example (needs oracle/_ref/fixtures at test time)
    with tempfile.TemporaryDirectory() as d:
        for f in os.listdir(FIXTURES):
            with open(os.path.join(FIXTURES, f), "rb") as src, open(os.path.join(d, f), "wb") as dst:
                dst.write(src.read())
        for mode in (0, 1, 2):
            g["listeria"][str(mode)] = stream_digests(d, "seqlist.txt", mode)
    # 4. hash known answers through the reference's table: position of a k-mer = the slot the reference wrote
    rng = np.random.default_rng(5)
    text = np.frombuffer(b"ACGT", dtype=np.uint8)[rng.integers(0, 4, 16 * 40 + 64)]
    r = RefMatcher(1 << 20, 32, 16, 1, 16)
    r.disable_sliding_window()
    r.load_ref(text, rc=False)
    tab = r.table()
    nz = np.nonzero(tab)[0]
    g["hash"] = {"text_seed": 5, "text_len": int(text.size), "slots": [[int(s), int(tab[s])] for s in nz]}
    r.close()
    g["coders"] = coders_golden()
    g["rcmatch"] = rcmatch_golden()
    with open(os.path.join(HERE, "golden.json"), "w") as f:
        json.dump(g, f, indent=1, sort_keys=True)
    print("wrote", os.path.join(HERE, "golden.json"))


if __name__ == "__main__":
    main()
