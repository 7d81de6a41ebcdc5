"""Seeded synthetic bacteria-like collections: a random ACGT base genome and independent
mutants of it (SNPs at Binomial(L, p_snp) uniform positions, alternating deletions / insertions of length
U[1,19] at Binomial(L, p_indel) sites).  Used by bench.py and the tests; plain numpy."""
import numpy as np

_ACGT = np.frombuffer(b"ACGT", dtype=np.uint8)


def synth_base(length, seed):
    rng = np.random.default_rng(seed)
    return _ACGT[rng.integers(0, 4, length)]


def synth_mutant(base, seed, p_snp, p_indel):
    rng = np.random.default_rng(seed)
    L = base.size
    g = base.copy()
    ns = rng.binomial(L, p_snp)
    if ns:
        pos = rng.integers(0, L, ns)
        code = np.searchsorted(_ACGT, g[pos])  # A < C < G < T in ASCII
        g[pos] = _ACGT[(code + rng.integers(1, 4, ns)) % 4]
    ni = rng.binomial(L, p_indel)
    if ni:
        sites = np.sort(rng.integers(0, L, ni))
        lens = rng.integers(1, 20, ni)
        parts, prev = [], 0
        for t, (sp, ln) in enumerate(zip(sites, lens)):
            if sp < prev:
                continue
            parts.append(g[prev:sp])
            if t % 2 == 0:
                prev = min(L, sp + ln)  # deletion
            else:
                parts.append(_ACGT[rng.integers(0, 4, ln)])  # insertion
                prev = sp
        parts.append(g[prev:])
        g = np.concatenate(parts)
    return g


def synth_collection(n, length, seed, p_snp, p_indel, first=1):
    """genome 0 = base; genomes first..first+n-2 = independent mutants (the index selects the mutant seed, so
    ranks of a sharded run draw disjoint genomes of one collection)."""
    base = synth_base(length, seed)
    return [base] + [synth_mutant(base, seed * 1000003 + i, p_snp, p_indel) for i in range(first, first + n - 1)]


def fasta_bytes(seq_len, header_len, width=80):
    """size of the FASTA record a sequence of seq_len bases comes from ('>' header '\\n' + 80-column lines)"""
    return 1 + header_len + 1 + seq_len + (seq_len + width - 1) // width


def write_fasta(path, header, seq, width=80):
    seq = np.ascontiguousarray(seq, dtype=np.uint8)
    n = seq.size
    full = n // width
    body = np.empty(n + (n + width - 1) // width, dtype=np.uint8)
    if full:
        blk = body[:full * (width + 1)].reshape(full, width + 1)
        blk[:, :width] = seq[:full * width].reshape(full, width)
        blk[:, width] = 10
    if n % width:
        body[full * (width + 1):-1] = seq[full * width:]
        body[-1] = 10
    with open(path, "wb") as f:
        f.write(b">" + header.encode() + b"\n")
        f.write(body.tobytes())


def mgmp_ref_limit(n_files, basic_len, rc=True):
    """reference-buffer limit MBGC would choose: MultipleGenomeMatchingProcessor.cpp:130-134,152-168
    (sub-4 GiB case): factor = 2^clamp(15 - clz32(files)/3, 5, 12), x max(|G0|, 2 MiB), x 2 with RC."""
    clz = 32 - int(n_files).bit_length()
    tmp = 15 - clz // 3
    tmp = 5 if tmp < 5 else (12 if tmp > 12 else tmp)
    lim = (1 << tmp) * max(basic_len, 1 << 21) * (2 if rc else 1)
    return min(lim, 2**32 - 1)


def _write_one(args):
    d, i, length, seed, p_snp, p_indel = args
    base = synth_base(length, seed)
    g = base if i == 0 else synth_mutant(base, seed * 1000003 + i, p_snp, p_indel)
    name = f"g{i:04d}.fna"
    path = f"{d}/{name}"
    write_fasta(path, f"syn_{i:04d} synthetic E. coli-like genome {i}", g)
    import os
    return name, os.path.getsize(path)


def write_collection(d, n, length, seed, p_snp, p_indel, workers=None):
    """the collection of SURVEY.md 8d as files: g0000.fna = base, g%04d.fna = mutant i (80-column lines, one record per
    file), plus list.txt in index order.  Returns (names, total FASTA bytes).  Files that already exist with the right
    count are kept (the generator is deterministic)."""
    import os
    from concurrent.futures import ProcessPoolExecutor
    os.makedirs(d, exist_ok=True)
    stamp = os.path.join(d, f".done_{n}_{length}_{seed}_{p_snp}_{p_indel}")
    names = [f"g{i:04d}.fna" for i in range(n)]
    if not os.path.exists(stamp):
        workers = workers or min(32, os.cpu_count() or 1)
        jobs = [(d, i, length, seed, p_snp, p_indel) for i in range(n)]
        if workers > 1 and n > 4:
            with ProcessPoolExecutor(workers) as ex:
                list(ex.map(_write_one, jobs, chunksize=max(1, n // (workers * 4))))
        else:
            for j in jobs:
                _write_one(j)
        with open(os.path.join(d, "list.txt"), "w") as f:
            f.write("\n".join(names) + "\n")
        open(stamp, "w").close()
    total = sum(os.path.getsize(os.path.join(d, x)) for x in names)
    return names, total
