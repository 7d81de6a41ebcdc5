#!/usr/bin/env python
"""Whole-pipeline compression through the reference's own CLI: stock `mbgc c` (oracle/_ref/mbgc) beside the batched drop-in
(oracle/_ref/mbgc_b200_batched = reference CLI + matcher shim + packer shim + the collection pass instead of the
sequential target loop) on a synthetic collection of SURVEY.md 8d's shape.

  stock, default threads   `mbgc c -m M`            what a user of the reference runs, 3 runs
  stock, -t 1              the deterministic schedule: the parity oracle for the archive bytes
  batched, -t 1            must write the SAME BYTES as stock -t 1
  batched, -Q              sequential schedule on the GPUs + all host threads for the back end: the timed GPU arm
                           (its archive equals the -t 1 one except the two thread-count bytes of the header)

    python tools/whole_pipeline.py --config c2near|c2div|c3|c4 [--gpus 1,2] [--dir /dev/shm/mbgc] [--genomes N]
"""
import argparse
import hashlib
import json
import os
import re
import subprocess
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from mbgc_b200.synth import write_collection  # noqa: E402

CONFIGS = {  # name: genomes, length, seed, p_snp, p_indel, mode
    "c2near": (100, 5_000_000, 1234, 1e-4, 1e-5, 0),
    "c2div": (100, 5_000_000, 1234, 1e-2, 1e-4, 1),
    "c3": (1000, 4_900_000, 2025, 1e-4, 1e-5, 1),
    "c4": (1000, 4_900_000, 2025, 1e-4, 1e-5, 3),
}
REF = os.path.join(ROOT, "oracle", "_ref")


def run(binary, args, cwd, env_extra=None):
    env = dict(os.environ)
    env.pop("OMP_NUM_THREADS", None)
    env.update(env_extra or {})
    t0 = time.perf_counter()
    r = subprocess.run([os.path.join(REF, binary)] + args, cwd=cwd, capture_output=True, text=True, env=env)
    dt = time.perf_counter() - t0
    if r.returncode != 0:
        print(r.stdout[-3000:], r.stderr[-3000:])
        raise SystemExit(f"{binary} {' '.join(args)} failed")
    log = r.stdout + r.stderr
    def ms(pat):
        m = re.search(pat + r" - (\d+) \[ms\]", log)
        return int(m.group(1)) if m else None
    out = {"wall_s": round(dt, 3), "loaded_reference_ms": ms("processed reference dataset"), "matching_finished_ms": ms("matching finished"),
           "encoding_time_ms": ms("encoding time")}
    tr = [l for l in log.splitlines() if l.startswith("[mbgc_b200] batched pass")]
    if tr:
        out["pass"] = tr[-1][len("[mbgc_b200] "):]
    cr = [l for l in log.splitlines() if "create: CUDA context" in l]
    if cr:
        out["cuda_context_ms"] = float(cr[0].split(" at ")[1].split()[0])
    return out


def md5(p):
    return hashlib.md5(open(p, "rb").read()).hexdigest()


def decodes(archive, d, names):
    """the stock decoder (sequential, `d -Q`) restores every file byte for byte"""
    import shutil
    out = os.path.join(d, "_decoded")
    shutil.rmtree(out, ignore_errors=True)
    os.makedirs(out)
    run("mbgc", ["d", "-Q", archive, out], d)
    ok = all(open(os.path.join(d, n), "rb").read() == open(os.path.join(out, n), "rb").read() for n in names)
    shutil.rmtree(out, ignore_errors=True)
    return ok


def measure(d, total, mode, gpus, stock_runs=3, skip_t1=False, parallel=False, names=None):
    """d: directory with the FASTA files and list.txt.  Returns the report dictionary."""
    out = {"fasta_bytes": total, "mode": mode, "host_cores": os.cpu_count()}
    m = ["c", "-m", str(mode), "-f"]
    arch = lambda tag: os.path.join(d, f"{tag}.mbgc")
    out["stock_default_threads"] = [run("mbgc", m + ["list.txt", arch("stock")], d) for _ in range(stock_runs)]
    if not skip_t1:
        out["stock_t1"] = run("mbgc", m + ["-t", "1", "list.txt", arch("stock_t1")], d)
        out["stock_t1"]["md5"] = md5(arch("stock_t1"))
    for g in gpus:
        env = {"MBGC_B200_DEVICES": ",".join(str(i) for i in range(g)), "MBGC_B200_TRACE": "1"}
        key = f"batched_{g}gpu"
        out[key] = {}
        if not skip_t1:
            r = run("mbgc_b200_batched", m + ["-t", "1", "list.txt", arch(key + "_t1")], d, env)
            r["md5"] = md5(arch(key + "_t1"))
            r["byte_identical_to_stock_t1"] = r["md5"] == out["stock_t1"]["md5"]
            out[key]["t1"] = r
        runs = [run("mbgc_b200_batched", m + ["-Q", "list.txt", arch(key + "_Q")], d, env) for _ in range(3)]
        out[key]["Q_default_threads"] = runs
        if not skip_t1:  # same archive but for the thread-count bytes of the header (1-based bytes 10, 11)
            x, y = open(arch(key + "_Q"), "rb").read(), open(arch(key + "_t1"), "rb").read()
            diff = [i for i in range(min(len(x), len(y))) if x[i] != y[i]] if len(x) == len(y) else None
            out[key]["Q_vs_t1_differing_bytes"] = diff if diff is None or len(diff) < 8 else len(diff)
    if parallel:  # the drop-in in the reference's default (parallel) mode: the pass's batch-frozen schedule
        out["stock_default_threads"][-1]["archive_bytes"] = os.path.getsize(arch("stock"))
        for g in gpus:
            env = {"MBGC_B200_DEVICES": ",".join(str(i) for i in range(g)), "MBGC_B200_TRACE": "1"}
            key = f"batched_{g}gpu"
            runs = [run("mbgc_b200_batched", m + ["list.txt", arch(key + "_par")], d, env) for _ in range(2)]
            out[key]["default_mode"] = {"runs": runs, "archive_bytes": os.path.getsize(arch(key + "_par")),
                                        "md5": md5(arch(key + "_par"))}
            if names:
                out[key]["default_mode"]["decodes_with_stock_decoder"] = decodes(arch(key + "_par"), d, names)
    best = min(out["stock_default_threads"], key=lambda r: r["wall_s"])
    out["summary"] = {"cpu_cli": {"what": "stock `mbgc c` with default threads, best of the runs", "wall_s": best["wall_s"],
                                  "GB_per_s": round(total / best["wall_s"] / 1e9, 3),
                                  "match_phase_ms": (best["matching_finished_ms"] or 0) - (best["loaded_reference_ms"] or 0),
                                  "encoding_time_ms": best["encoding_time_ms"], "cores": os.cpu_count()}}
    for g in gpus:
        b = min(out[f"batched_{g}gpu"]["Q_default_threads"], key=lambda r: r["wall_s"])
        ctx = b.get("cuda_context_ms") or 0.0
        out["summary"][f"batched_{g}gpu"] = {
            "wall_s": b["wall_s"], "GB_per_s": round(total / b["wall_s"] / 1e9, 3),
            "match_phase_ms": (b["matching_finished_ms"] or 0) - (b["loaded_reference_ms"] or 0),
            "encoding_time_ms": b["encoding_time_ms"], "cuda_context_ms": ctx,
            "encoding_time_without_context_ms": (b["encoding_time_ms"] or 0) - ctx}
        if parallel:
            p = min(out[f"batched_{g}gpu"]["default_mode"]["runs"], key=lambda r: r["wall_s"])
            out["summary"][f"batched_{g}gpu"]["default_mode"] = {
                "wall_s": p["wall_s"], "match_phase_ms": (p["matching_finished_ms"] or 0) - (p["loaded_reference_ms"] or 0),
                "encoding_time_ms": p["encoding_time_ms"], "cuda_context_ms": p.get("cuda_context_ms")}
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--config", default="c2near", choices=list(CONFIGS))
    ap.add_argument("--genomes", type=int, default=0)
    ap.add_argument("--gpus", default="1")
    ap.add_argument("--dir", default="/dev/shm/mbgc_wp")
    ap.add_argument("--stock-runs", type=int, default=3)
    ap.add_argument("--skip-t1", action="store_true", help="skip the two -t 1 runs (archive identity)")
    ap.add_argument("--parallel", action="store_true", help="also run the drop-in in the default (parallel) mode and decode its archive")
    a = ap.parse_args()
    n, length, seed, p_snp, p_indel, mode = CONFIGS[a.config]
    if a.genomes:
        n = a.genomes
    d = os.path.join(a.dir, f"{a.config}_{n}")
    t0 = time.perf_counter()
    names, total = write_collection(d, n, length, seed, p_snp, p_indel)
    out = {"config": a.config, "genomes": n, "generate_s": round(time.perf_counter() - t0, 1)}
    out.update(measure(d, total, mode, [int(x) for x in a.gpus.split(",")], a.stock_runs, a.skip_t1, a.parallel, names))
    print(json.dumps(out, indent=1))
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    with open(os.path.join(ROOT, "gpurun_out", f"whole_pipeline_{a.config}_{n}.json"), "w") as f:
        json.dump(out, f, indent=1)


if __name__ == "__main__":
    main()
