#!/usr/bin/env python
"""Whole-pipeline DEcompression through the reference's own CLI: stock `mbgc d` (oracle/_ref/mbgc) beside the decoder drop-in
(oracle/_ref/mbgc_b200_decode = reference CLI + the target loop of MBGC_Decoder on the device) on an archive the stock
encoder wrote; every extracted file is compared with the original.

    python tools/whole_pipeline_decode.py --config c2near|c2div|c3 [--dir /dev/shm/mbgc_dec] [--genomes N]
"""
import argparse
import filecmp
import json
import os
import re
import shutil
import subprocess
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from mbgc_b200.synth import write_collection  # noqa: E402

CONFIGS = {"c2near": (100, 5_000_000, 1234, 1e-4, 1e-5, 0), "c2div": (100, 5_000_000, 1234, 1e-2, 1e-4, 1),
           "c3": (1000, 4_900_000, 2025, 1e-4, 1e-5, 1)}
REF = os.path.join(ROOT, "oracle", "_ref")


def run(binary, args, cwd, env_extra=None):
    env = dict(os.environ)
    env.pop("OMP_NUM_THREADS", None)
    env.update(env_extra or {})
    t0 = time.perf_counter()
    r = subprocess.run([os.path.join(REF, binary)] + args, cwd=cwd, capture_output=True, text=True, env=env)
    dt = time.perf_counter() - t0
    if r.returncode != 0:
        print(r.stdout[-2000:], r.stderr[-2000:])
        raise SystemExit(f"{binary} {' '.join(args)} failed")
    log = r.stdout + r.stderr
    out = {"wall_s": round(dt, 3)}
    m = re.search(r"loaded streams - (\d+) \[ms\]", log)
    if m:
        out["loaded_streams_ms"] = int(m.group(1))
    tr = [l for l in log.splitlines() if l.startswith("[mbgc_b200] decode:")]
    if tr:
        out["device"] = tr[-1][len("[mbgc_b200] "):]
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--config", default="c2near", choices=sorted(CONFIGS))
    ap.add_argument("--dir", default="/dev/shm/mbgc_dec")
    ap.add_argument("--genomes", type=int, default=0)
    a = ap.parse_args()
    n, length, seed, ps, pi, mode = CONFIGS[a.config]
    if a.genomes:
        n = a.genomes
    d = os.path.join(a.dir, a.config)
    names, total = write_collection(d, n, length, seed, ps, pi)
    with open(os.path.join(d, "list.txt"), "w") as f:
        f.write("\n".join(names) + "\n")
    arch = os.path.join(d, "stock.mbgc")
    enc = run("mbgc", ["c", "-m", str(mode), "-f", "list.txt", arch], d)
    res = {"config": a.config, "genomes": n, "fasta_bytes": total, "archive_bytes": os.path.getsize(arch), "host_cores": os.cpu_count(),
           "encode_stock": enc, "decode": {}}
    for tag, binary, args, env in (("stock_default_threads", "mbgc", [], None), ("stock_Q", "mbgc", ["-Q"], None),
                                   ("device_default_threads", "mbgc_b200_decode", [], {"MBGC_B200_TRACE": "1"}),
                                   ("device_second_run", "mbgc_b200_decode", [], {"MBGC_B200_TRACE": "1"})):
        out = os.path.join(d, "out")
        shutil.rmtree(out, ignore_errors=True)
        os.makedirs(out)
        r = run(binary, ["d"] + args + [arch, out], d, env)
        r["GB_per_s"] = round(total / r["wall_s"] / 1e9, 3)
        r["all_files_identical"] = all(filecmp.cmp(os.path.join(d, x), os.path.join(out, x), shallow=False) for x in names)
        res["decode"][tag] = r
    shutil.rmtree(os.path.join(d, "out"), ignore_errors=True)
    print(json.dumps(res))


if __name__ == "__main__":
    main()
