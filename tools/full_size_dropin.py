#!/usr/bin/env python
"""Full-size drop-in check (BASELINE.json configs[1] shape): the reference CLI with our matcher linked in
(oracle/_ref/mbgc_b200; and oracle/_ref/mbgc_b200_full with the stream packer on the GPU as well) against the stock CLI (oracle/_ref/mbgc), both `c -t 1 -m M`, on the synthetic
100 x 5 Mbp collection.  Prints archive md5s (must be equal), wall times and the reference's own timing lines.

    python tools/full_size_dropin.py [--genomes 100] [--length 5000000] [--tier near|div] [--mode 0] [--dir /tmp/c2]
"""
import argparse
import hashlib
import json
import os
import subprocess
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from mbgc_b200.synth import synth_base, synth_mutant, write_fasta  # noqa: E402

TIERS = {"near": (1e-4, 1e-5), "div": (1e-2, 1e-4)}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--genomes", type=int, default=100)
    ap.add_argument("--length", type=int, default=5_000_000)
    ap.add_argument("--tier", default="near", choices=list(TIERS))
    ap.add_argument("--mode", type=int, default=0)
    ap.add_argument("--dir", default="/tmp/mbgc_c2")
    a = ap.parse_args()
    os.makedirs(a.dir, exist_ok=True)
    p_snp, p_indel = TIERS[a.tier]
    base = synth_base(a.length, 1234)
    names = []
    total = 0
    for i in range(a.genomes):
        g = base if i == 0 else synth_mutant(base, 1234 * 1000003 + i, p_snp, p_indel)
        name = f"g{i:04d}.fna"
        write_fasta(os.path.join(a.dir, name), f"syn_{i:04d} synthetic E. coli-like genome {i}", g)
        total += os.path.getsize(os.path.join(a.dir, name))
        names.append(name)
    with open(os.path.join(a.dir, "list.txt"), "w") as f:
        f.write("\n".join(names) + "\n")
    env = dict(os.environ)
    env.pop("OMP_NUM_THREADS", None)
    out = {"fasta_bytes": total, "genomes": a.genomes, "tier": a.tier, "mode": a.mode}
    # mbgc_b200 = our matcher behind the reference CLI; mbgc_b200_full = matcher and MBGC_Encoder::processMatches (K6)
    for tag, binary in (("reference", "mbgc"), ("b200_dropin", "mbgc_b200"), ("b200_dropin_full", "mbgc_b200_full")):
        arch = os.path.join(a.dir, f"{tag}.mbgc")
        if os.path.exists(arch):
            os.remove(arch)
        t0 = time.perf_counter()
        r = subprocess.run([os.path.join(ROOT, "oracle", "_ref", binary), "c", "-t", "1", "-m", str(a.mode), "-f", "list.txt", arch],
                           cwd=a.dir, capture_output=True, text=True, env=env)
        dt = time.perf_counter() - t0
        if r.returncode != 0:
            print(r.stdout[-2000:], r.stderr[-2000:])
            raise SystemExit(f"{binary} failed")
        timing = [l.strip() for l in (r.stdout + r.stderr).splitlines() if "time" in l.lower() or "finished" in l.lower()]
        out[tag] = {"wall_s": dt, "GB_per_s": total / dt / 1e9, "archive_bytes": os.path.getsize(arch),
                    "md5": hashlib.md5(open(arch, "rb").read()).hexdigest(), "log": timing[-4:]}
    out["byte_identical"] = out["reference"]["md5"] == out["b200_dropin"]["md5"] == out["b200_dropin_full"]["md5"]
    print(json.dumps(out, indent=1))
    if not out["byte_identical"]:
        raise SystemExit("archives differ")


if __name__ == "__main__":
    main()
