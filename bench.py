#!/usr/bin/env python
"""bench.py -- MBGC match-stage throughput on B200 (and the reference's CPU path beside it).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P \
        bench.py --gpus N --steps K --warmup W

Workload (BASELINE.json configs[1]): a synthetic 100-genome E. coli-like collection -- a random 5 Mbp base
genome (G0) and independent SNP/indel mutants of it ("near" tier of SURVEY.md 8d: p_snp 1e-4, p_indel 1e-5).
The matcher state is what MBGC has after loading G0 (+ its reverse complement); one STEP = one pass of the hot
path (probe -> seed runs -> extension -> sequential-equivalent resolve -> match lists on the host) over the
rank's batch of --genomes target genomes against that frozen state.  With N ranks every rank matches its own
--genomes mutants of the same collection against its replica of the index (weak scaling, no data-path
collective: SURVEY.md 8e).

Metric: compression input GB/s of the match stage = FASTA bytes of the matched genomes / time.
  value  device-resident inputs, CUDA-event time on the matcher's stream, max over ranks
  e2e    same call through the C ABI with pinned HOST buffers (H2D of the texts and D2H of the match lists
         inside the timed region)
  roofline  dominant kernel (k_probe): 6.0 algorithmic bytes per target position / its mean
         launch time, against the measured HBM copy bandwidth of MEASURED_PEAKS.json
  cpu_baseline  the reference's own matcher (oracle/_ref/libref_matcher.so; else the C oracle port) on the
         host cores over the same batch
"""
import argparse
import ctypes as C
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

from mbgc_b200.synth import fasta_bytes, mgmp_ref_limit, synth_base, synth_mutant  # noqa: E402

SEED = 1234
HEADER_LEN = len("syn_0000 synthetic E. coli-like genome 0")
TIERS = {"near": (1e-4, 1e-5), "div": (1e-2, 1e-4)}
ALG_BYTES_PER_POSITION = 6.0  # 1 ASCII + .25 + .25 packed + 4 table probe + .25 ref + .25 out


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--genomes", type=int, default=100, help="target genomes per rank")
    ap.add_argument("--length", type=int, default=5_000_000)
    ap.add_argument("--tier", default="near", choices=list(TIERS))
    ap.add_argument("--mode", type=int, default=0, help="MBGC mode the matcher parameters come from (0..3)")
    ap.add_argument("--cpu-sample", type=int, default=100, help="genomes of the batch the CPU baseline is timed on")
    ap.add_argument("--cpu-sample-other", type=int, default=32, help="same for the secondary tier")
    ap.add_argument("--ingest-genomes", type=int, default=100, help="genomes in the multi-FASTA image of the ingest measurement")
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--batches", type=int, default=2, help="distinct batches the timed steps rotate through")
    ap.add_argument("--long-seconds", type=float, default=1.0, help="length of the additional long timed region")
    ap.add_argument("--no-decode", action="store_true", help="skip the decoder leg (row f4: collection pass -> streams -> device decode)")
    ap.add_argument("--no-c4", action="store_true", help="skip the mode-3 run of the collection (C4: C3 + the reverse-complement pass)")
    ap.add_argument("--no-collection", action="store_true", help="skip the C3 collection pass (1k genomes, strong scaling)")
    ap.add_argument("--collection-genomes", type=int, default=1000)
    ap.add_argument("--collection-length", type=int, default=4_900_000)
    ap.add_argument("--no-cli", action="store_true", help="skip the whole-pipeline CLI runs (stock mbgc vs the batched drop-in)")
    ap.add_argument("--scratch", default="/dev/shm" if os.path.isdir("/dev/shm") else "/tmp")
    ap.add_argument("--no-other-tier", action="store_true", help="skip the secondary similarity tier")
    return ap.parse_args()


def make_batch(args, rank, batch=0):
    """G0 and this rank's target genomes (disjoint mutant indices per rank and per rotating batch)."""
    p_snp, p_indel = TIERS[args.tier]
    base = synth_base(args.length, SEED)
    first = 1 + (rank * max(1, args.batches) + batch) * args.genomes
    targets = [synth_mutant(base, SEED * 1000003 + i, p_snp, p_indel) for i in range(first, first + args.genomes)]
    return base, targets


def skip_margin_of(mode):
    return 24 if mode >= 2 else 16  # MGMP_Params.h:33-34, MBGC_Params.h:906-909


def peaks():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md)"


def probe_traffic():
    """measured DRAM bytes per target position of the main kernels (profiles/traffic.json, written from the
    committed `ncu --set full` captures of this workload)"""
    try:
        with open(os.path.join(ROOT, "profiles", "traffic.json")) as f:
            return json.load(f)
    except Exception:
        return None


class ClockSampler(threading.Thread):
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu):
        super().__init__(daemon=True)
        self.gpu = gpu
        self.samples = []
        self.stop_flag = False

    def run(self):
        while not self.stop_flag:
            try:
                out = subprocess.run(["nvidia-smi", "-i", str(self.gpu), "--query-gpu=" + self.Q,
                                      "--format=csv,noheader,nounits"], capture_output=True, text=True, timeout=5).stdout
                f = [x.strip() for x in out.strip().split(",")]
                if len(f) >= 8:
                    self.samples.append(f)
            except Exception:
                pass
            time.sleep(0.1)

    def summary(self):
        if not self.samples:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["unsampled"]}
        sm = sorted(float(s[1]) for s in self.samples)
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = [n for k, n in enumerate(names) if any(s[4 + k] == "Active" for s in self.samples)]
        return {"sm_mhz": sm[len(sm) // 2], "sm_max_mhz": float(self.samples[0][2]), "reasons": reasons,
                "samples": len(self.samples)}


# ------------------------------------------------------------------------------------------------------
# CPU arm: the reference's own matcher on the host cores
# ------------------------------------------------------------------------------------------------------

def cpu_match_stage(base, targets, limit, skip_margin, threads):
    """Returns (seconds, kind, total matches).  kind = "reference" when oracle/_ref/libref_matcher.so (the
    reference's SlidingWindowSparseEMMatcher.cpp built unmodified) is there, else "port" (oracle C restatement,
    single thread)."""
    ref_so = os.path.join(ROOT, "oracle", "_ref", "libref_matcher.so")
    n = len(targets)
    ptrs = (C.c_void_p * n)(*[t.ctypes.data for t in targets])
    lens = (C.c_uint64 * n)(*[t.size for t in targets])
    if os.path.exists(ref_so):
        os.environ.setdefault("OMP_NUM_THREADS", "1")  # index build is racy with more
        L = C.CDLL(ref_so)
        L.refm_create.restype = C.c_void_p
        L.refm_create.argtypes = [C.c_uint64, C.c_uint32, C.c_int, C.c_int, C.c_int]
        L.refm_disable_sliding_window.argtypes = [C.c_void_p]
        L.refm_load_ref.argtypes = [C.c_void_p, C.c_void_p, C.c_uint64, C.c_int, C.c_int, C.c_char]
        L.refm_match_many.restype = C.c_uint64
        L.refm_match_many.argtypes = [C.c_void_p, C.c_int, C.POINTER(C.c_void_p), C.POINTER(C.c_uint64), C.c_uint32,
                                      C.c_int, C.c_void_p]
        L.refm_destroy.argtypes = [C.c_void_p]
        h = L.refm_create(limit, 32, 16, 1, skip_margin)
        L.refm_disable_sliding_window(h)
        L.refm_load_ref(h, base.ctypes.data, base.size, 1, 1, b"\0")
        t0 = time.perf_counter()
        total = L.refm_match_many(h, n, ptrs, lens, 32, threads, None)
        dt = time.perf_counter() - t0
        L.refm_destroy(h)
        return dt, "reference", int(total), threads
    subprocess.run(["make", "-s", "oracle"], cwd=os.path.join(ROOT, "oracle"), check=True)
    L = C.CDLL(os.path.join(ROOT, "oracle", "libmbgc_oracle.so"))
    L.mbo_create.restype = C.c_void_p
    L.mbo_create.argtypes = [C.c_uint64, C.c_uint32, C.c_int, C.c_int, C.c_int]
    L.mbo_disable_sliding_window.argtypes = [C.c_void_p]
    L.mbo_load_ref.argtypes = [C.c_void_p, C.c_void_p, C.c_uint64, C.c_int, C.c_int, C.c_uint8]
    L.mbo_match.restype = C.c_uint64
    L.mbo_match.argtypes = [C.c_void_p, C.c_void_p, C.c_uint64, C.c_uint32, C.c_uint64, C.c_int]
    L.mbo_destroy.argtypes = [C.c_void_p]
    h = L.mbo_create(limit, 32, 16, 1, skip_margin)
    L.mbo_disable_sliding_window(h)
    L.mbo_load_ref(h, base.ctypes.data, base.size, 1, 1, 0)
    t0 = time.perf_counter()
    total = 0
    for t in targets:
        total += L.mbo_match(h, t.ctypes.data, t.size, 32, 2**64 - 1, 1)
    dt = time.perf_counter() - t0
    L.mbo_destroy(h)
    return dt, "port", int(total), 1


def cpu_pack_stage(base, targets, lists, limit, skip_margin, mode):
    """MBGC_Encoder::processMatches on one host core through the oracle port (oracle/mbgc_oracle.c:
    mbo_process_matches; the reference's own function is a private member of its encoder and not reachable from
    the matcher driver).  Returns (seconds, stream bytes)."""
    subprocess.run(["make", "-s", "oracle"], cwd=os.path.join(ROOT, "oracle"), check=True)
    L = C.CDLL(os.path.join(ROOT, "oracle", "libmbgc_oracle.so"))
    L.mbo_create.restype = C.c_void_p
    L.mbo_create.argtypes = [C.c_uint64, C.c_uint32, C.c_int, C.c_int, C.c_int]
    L.mbo_disable_sliding_window.argtypes = [C.c_void_p]
    L.mbo_load_ref.argtypes = [C.c_void_p, C.c_void_p, C.c_uint64, C.c_int, C.c_int, C.c_uint8]
    L.mbo_loading_position.restype = C.c_uint64
    L.mbo_loading_position.argtypes = [C.c_void_p]
    L.mbo_destroy.argtypes = [C.c_void_p]
    L.mbo_streams_create.restype = C.c_void_p
    L.mbo_streams_destroy.argtypes = [C.c_void_p]
    L.mbo_stream_size.restype = C.c_uint64
    L.mbo_stream_size.argtypes = [C.c_void_p, C.c_int]
    L.mbo_process_matches.restype = C.c_uint64
    L.mbo_process_matches.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.POINTER(C.c_uint64), C.c_void_p,
                                      C.c_uint64, C.c_uint64, C.POINTER(C.c_uint64), C.c_uint64]
    h = L.mbo_create(limit, 32, 16, 1, skip_margin)
    L.mbo_disable_sliding_window(h)
    L.mbo_load_ref(h, base.ctypes.data, base.size, 1, 1, 0)
    # mbo_pack_params_t (oracle/mbgc_oracle.h): 14 ints, defaults of MBGC_Params.h:66-97
    pp = (C.c_int * 14)(1, 256, 64, 2, 1, 0 if mode == 0 else 1, 0, 1, 125, 500, 50, 50, 32, 128)
    loaded = (C.c_uint64 * 1)(L.mbo_loading_position(h))
    st = L.mbo_streams_create()
    copies = [np.ascontiguousarray(m, dtype=np.uint64).copy() for m in lists]  # the call edits its list in place
    t0 = time.perf_counter()
    for t, m in zip(targets, copies):
        n = C.c_uint64(len(m))
        L.mbo_process_matches(st, pp, h, m.ctypes.data if len(m) else None, C.byref(n), t.ctypes.data, t.size,
                              2**64 - 1, loaded, 1)
    dt = time.perf_counter() - t0
    nbytes = sum(L.mbo_stream_size(st, i) for i in range(6))
    L.mbo_streams_destroy(st)
    L.mbo_destroy(h)
    return dt, int(nbytes)


def fasta_image(targets, first_index):
    """the targets as one multi-FASTA file image (80-column lines, the headers of SURVEY.md 8d)"""
    parts = []
    for k, t in enumerate(targets):
        i = first_index + k
        parts.append(np.frombuffer(f">syn_{i:04d} synthetic E. coli-like genome {i}\n".encode(), dtype=np.uint8))
        full = t.size // 80 * 80
        body = np.empty((full // 80, 81), dtype=np.uint8)
        body[:, :80] = t[:full].reshape(-1, 80)
        body[:, 80] = 10
        parts.append(body.reshape(-1))
        if t.size > full:
            parts.append(np.concatenate([t[full:], np.array([10], dtype=np.uint8)]))
    return np.concatenate(parts)


def cpu_fasta_reader(image):
    """the reference's own reader (utils/kseq.h through oracle/_ref/libref_kseq.so; else the oracle port) on one host
    core over the same file image.  Returns (seconds, kind, records)."""
    n = image.size
    ref_so = os.path.join(ROOT, "oracle", "_ref", "libref_kseq.so")
    cap = 4096
    recs = np.zeros((cap, 4), dtype=np.uint64)
    seq = np.empty(n, dtype=np.uint8)
    dna, mx, st = C.c_uint64(), C.c_uint64(), C.c_int()
    if os.path.exists(ref_so):
        L = C.CDLL(ref_so)
        L.refk_fasta_parse.restype = C.c_uint64
        L.refk_fasta_parse.argtypes = [C.c_void_p, C.c_uint64, C.c_void_p, C.c_uint64, C.c_void_p, C.c_void_p,
                                       C.POINTER(C.c_uint64), C.POINTER(C.c_uint64), C.POINTER(C.c_int)]
        hdr = np.empty(1 << 20, dtype=np.uint8)
        t0 = time.perf_counter()
        k = L.refk_fasta_parse(image.ctypes.data, n, recs.ctypes.data, cap, seq.ctypes.data, hdr.ctypes.data,
                               C.byref(dna), C.byref(mx), C.byref(st))
        return time.perf_counter() - t0, "reference", int(k)
    subprocess.run(["make", "-s", "oracle"], cwd=os.path.join(ROOT, "oracle"), check=True)
    L = C.CDLL(os.path.join(ROOT, "oracle", "libmbgc_oracle.so"))
    L.mbo_fasta_parse.restype = C.c_uint64
    L.mbo_fasta_parse.argtypes = [C.c_void_p, C.c_uint64, C.c_void_p, C.c_uint64, C.c_void_p, C.POINTER(C.c_uint64),
                                  C.POINTER(C.c_uint64), C.POINTER(C.c_int)]
    t0 = time.perf_counter()
    k = L.mbo_fasta_parse(image.ctypes.data, n, recs.ctypes.data, cap, seq.ctypes.data, C.byref(dna), C.byref(mx), C.byref(st))
    return time.perf_counter() - t0, "port", int(k)


def bench_config(args, world):
    """the same dictionary in both arms (the driver compares them)"""
    return {"workload": workload_name(args), "genomes_per_gpu": args.genomes, "tier": args.tier, "mode": args.mode,
            "parallelism": f"targets sharded over {world} GPU(s), index replicated",
            "l2": "inputs (>= 500 MB per step, two batches rotating) exceed the 126 MB L2; no explicit flush"}


def workload_name(args):
    return (f"synthetic {args.genomes}-genome E. coli-like collection per GPU ({args.length / 1e6:g} Mbp base + "
            f"SNP/indel mutants, tier {args.tier}), mode {args.mode} matcher parameters")


def run_reference(args, rank):
    if rank != 0:
        return
    base, targets = make_batch(args, 0)
    sample = targets[:max(1, min(args.cpu_sample, len(targets)))]
    limit = mgmp_ref_limit(args.genomes * args.gpus + 1, base.size)
    threads = os.cpu_count() or 1
    fbytes = sum(fasta_bytes(t.size, HEADER_LEN) for t in sample)
    times = []
    kind, cores, total = "port", 1, 0
    for it in range(args.warmup + args.steps):
        dt, kind, total, cores = cpu_match_stage(base, sample, limit, skip_margin_of(args.mode), threads)
        if it >= args.warmup:
            times.append(dt)
    sec = sum(times) / len(times)
    v = fbytes / sec / 1e9
    line = {
        "impl": "reference", "metric": "compression input GB/s (match stage)", "value": v, "unit": "GB/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": sec * 1e3,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "u8", "data": "synthetic",
        "config": bench_config(args, args.gpus),
        "details": {"sample_genomes": len(sample), "matches": total},
        "cpu_baseline": {"value": v, "unit": "GB/s", "cores": cores, "kind": kind,
                         "sample": f"{len(sample)} of the {args.genomes} target genomes per step "
                                   f"({fbytes / 1e6:.0f} MB FASTA), matchTexts against the frozen G0+RC index"},
        "e2e": {"value": v, "unit": "GB/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


# Algorithmic bytes per target position of each kernel group (DESIGN.md section 4 derives them).
#   contract  SURVEY.md 8d: 6.0 B per target position for the whole match stage (probe every position)
#   k_probe   the same 6.0 B, for the positions it actually probes (all of them without the fast path)
#   k_clean   1 B text + (16 B descriptor + 8 B anchor + 16 B memo words + 13 B result) per 512-position chunk;
#             the reference bytes on the diagonal are a second byte per position, served by L2 while the loaded
#             reference fits it (10 MB here) and by HBM otherwise
#   k_anchor  48 B text + 16 occupancy words + table slot + 32 B reference + 24 B descriptor/result per chunk
STAGE_KERNELS = {"anchor": "k_anchor", "clean": "k_clean", "fullprobe": "k_probe",
                 "compact": "k_compact + k_heads + k_runs + scans", "extend": "k_extend + k_links",
                 "resolve": "k_island_* + k_resolve + k_gather"}


def kernel_alg_bytes(group, counters, ref_bytes):
    pos, chunks = counters["positions"], max(1, counters["chunks"])
    if group == "fullprobe":
        return ALG_BYTES_PER_POSITION * 16.0 * counters["lanes_full_probe"]
    if group == "clean":
        return pos * (1.0 + (0.0 if ref_bytes <= 100e6 else 1.0)) + chunks * 53.0
    if group == "anchor":
        return chunks * (48 + 64 + 8 + 32 + 24.0)
    return None


def stage_moved_bytes(counters, ref_bytes, n_matches):
    """bytes the match stage MUST move under the algorithm it actually runs (fast path + probe of the dirty lanes), per
    kernel group -- what `roofline.stage_moved` divides by the stage time.  Text-streaming kernels as in kernel_alg_bytes;
    the small kernels by their records: seed segments 16 B written + read twice (compact -> heads -> runs), the per-chunk
    counts / scans 12 B per chunk, runs 24 B written + read by extend, links, islands and resolve (4 x) + ~64 B of text /
    reference per extended run end, matches 16 B + 32 B written and 32 B read by the gather."""
    out = {g: kernel_alg_bytes(g, counters, ref_bytes) or 0.0 for g in ("anchor", "clean", "fullprobe")}
    out["compact"] = counters["seeds"] * 16.0 * 3 + counters["chunks"] * 12.0 + counters["runs"] * 24.0
    out["extend"] = counters["runs"] * (24.0 * 2 + 2 * 64.0)
    out["resolve"] = counters["runs"] * 24.0 * 2 + n_matches * (16.0 + 32.0 + 32.0)
    return out


def measure_tier(args, tier, rank, world, dist, local):
    """(line fragments for one tier) value / e2e / stage times / counters on this rank's batch"""
    import torch
    import mbgc_b200
    from mbgc_b200.dist import max_over_ranks

    targs = argparse.Namespace(**vars(args))
    targs.tier = tier
    nb = max(1, args.batches if tier == args.tier else 1)   # the timed steps rotate through nb distinct batches
    base, targets = make_batch(targs, rank, 0)
    batches = [targets] + [make_batch(targs, rank, b)[1] for b in range(1, nb)]
    n = len(targets)
    limit = mgmp_ref_limit(args.genomes * world + 1, base.size)
    sm = skip_margin_of(args.mode)
    m = mbgc_b200.Matcher(limit, 32, 16, 1, sm, device=local)
    m.disableSlidingWindow()
    t0 = time.perf_counter()
    m.loadRef(base, True, True, 0)
    index_s = time.perf_counter() - t0
    index_tm = m.loadRefTiming()
    # the same call once more on a scratch matcher whose staging buffers and mirror pages exist already (what every
    # reference extension after the first pays), with the host mirror on and off
    index_warm = {}
    if tier == args.tier:
        w = mbgc_b200.Matcher(8 * base.size + (64 << 20), 32, 16, 1, sm, device=local)
        w.disableSlidingWindow()
        w.loadRef(base, True, True, 0)
        w.loadRef(base, True, True, 0)
        index_warm["mirror_on"] = w.loadRefTiming()
        w.setHostMirror(False)
        w.loadRef(base, True, True, 0)
        index_warm["mirror_off"] = w.loadRefTiming()
        w.close()

    # device-resident batches: texts at 16-byte aligned offsets, 64-byte gaps, 8 KiB tail
    dev = []
    for tg in batches:
        offs, cur = [], 64
        for t in tg:
            offs.append(cur)
            cur = (cur + t.size + 64 + 15) // 16 * 16
        dbuf = torch.zeros(cur + 8192, dtype=torch.uint8, device="cuda")
        pinned = []
        for t, o in zip(tg, offs):
            ht = torch.from_numpy(t).pin_memory()
            pinned.append(ht)
            dbuf[o:o + t.size].copy_(ht, non_blocking=True)
        ls = [t.size for t in tg]
        dev.append(dict(dbuf=dbuf, pinned=pinned, offs=offs, lens=ls, offs_a=np.array(offs, dtype=np.uint64),
                        lens_a=np.array(ls, dtype=np.uint64), host=[p.numpy() for p in pinned],
                        fbytes=sum(fasta_bytes(x, HEADER_LEN) for x in ls)))
    torch.cuda.synchronize()
    dbuf, offs, lens, offs_a, lens_a, pinned = (dev[0][k] for k in ("dbuf", "offs", "lens", "offs_a", "lens_a", "pinned"))
    fbytes = dev[0]["fbytes"]
    stream = torch.cuda.ExternalStream(m.stream())

    def barrier():
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, steps, warmup, stage_runs=None):
        """fn(step index) -> result.  Returns (device ms of exactly `steps` calls, mean stage ms, launches, last result,
        FASTA bytes the timed steps covered)"""
        m.setStageTiming(False)  # the timed calls carry no stage events; the stage times come from the extra run below
        for k in range(warmup):
            fn(k)
        launches0 = m.totalLaunches()
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        res = None
        for k in range(steps):  # the timed region: exactly `steps` calls of the public API, nothing else
            res = fn(k)
        e1.record(stream)
        barrier()
        ms = max_over_ranks(e0.elapsed_time(e1), "cuda")  # device time, max over ranks
        launches = m.totalLaunches() - launches0
        covered = sum(dev[k % nb]["fbytes"] for k in range(steps))
        m.setStageTiming(True)
        # per-kernel-group times of the same call (events inside the library), collected outside the timed region
        stage = {k: 0.0 for k in mbgc_b200.matcher.T_NAMES}
        runs = stage_runs if stage_runs is not None else min(steps, 20)
        for k in range(runs):
            fn(k)
            for name, v in m.timing().items():
                stage[name] += v
        return ms, {k: v / max(1, runs) for k, v in stage.items()}, launches, res, covered

    def dev_step(k):
        b = dev[k % nb]
        return m.matchBatchDevice(b["dbuf"].data_ptr(), b["offs_a"], b["lens_a"], 32, copy=False, split=False)

    def host_step(k):
        return m.matchBatch(dev[k % nb]["host"], 32, copy=False, split=False)

    steps, warmup = args.steps, args.warmup
    ms_dev, stage_dev, launches, res, covered_dev = timed(dev_step, steps, warmup)
    # batch 0 alone, for the counters / roofline bytes (every batch has the same shape)
    res0 = m.matchBatchDevice(dbuf.data_ptr(), offs_a, lens_a, 32, copy=False, split=False)
    n_matches = int(len(res0[0]))
    counters = m.counters()
    ms_e2e, stage_e2e, _, res2, covered_e2e = timed(host_step, steps, warmup)
    res2 = m.matchBatch(dev[0]["host"], 32, copy=False, split=False)
    wire = m.wireBytes()
    assert n_matches == int(len(res2[0])), "device-resident and host-buffer passes disagree"
    # a long timed region (>= --long-seconds) of the same device-resident steps: the K-step number above is short
    long_run = None
    if tier == args.tier and args.long_seconds > 0:
        per = max(ms_dev / max(1, steps), 1e-3)
        lsteps = int(min(20000, max(steps, args.long_seconds * 1e3 / per + 1)))
        ms_long, _, _, _, covered_long = timed(dev_step, lsteps, 2, stage_runs=0)
        long_run = {"steps": lsteps, "seconds": ms_long / 1e3, "ms_per_step": ms_long / lsteps,
                    "value": covered_long * world / (ms_long / 1e3) / 1e9, "unit": "GB/s", "batches_rotated": nb}
    # K6: match stage + stream packing with everything resident (the lists never leave the device between the two)
    from mbgc_b200.matcher import PackParams
    pd = PackParams.for_mode(args.mode)
    loaded = [m.getLoadingPosition()]
    pack_tm = {}

    def match_and_pack():
        m.matchBatchDevice(dbuf.data_ptr(), offs_a, lens_a, 32, fetch=False)  # lists stay in HBM, only counts come back
        return m.packBatch(pd, None, None, loadedPosArr=loaded, lens=lens_a, raw=True)

    psteps = max(2, steps // 2)
    for _ in range(2):
        match_and_pack()
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    for _ in range(psteps):
        pres = match_and_pack()
        for k, v in m.packTiming().items():
            if k != "sweeps":
                pack_tm[k] = pack_tm.get(k, 0.0) + v / psteps
    e1.record(stream)
    barrier()
    ms_mp = max_over_ranks(e0.elapsed_time(e1), "cuda") / psteps
    pack = {"match_plus_pack_ms_per_step": ms_mp, "pack_ms": pack_tm, "sweeps": list(m.packTiming()["sweeps"]),
            "stream_bytes": pres["stream_bytes"],
            "kept_matches": int(pres["kept"].sum()), "unmatched_chars": int(pres["unmatched"].sum())}
    # FASTA ingest: the same genomes as one multi-FASTA file image in pinned host memory -> device texts
    ingest = None
    if tier == args.tier:
        ki = max(1, min(args.ingest_genomes, n))
        image = torch.from_numpy(fasta_image(targets[:ki], 1 + rank * args.genomes)).pin_memory()
        img = image.numpy()
        for _ in range(2):
            m.ingestFasta(img)
        isteps, tm = max(2, steps // 2), {}
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        for _ in range(isteps):
            ing = m.ingestFasta(img)
            for k, v in m.ingestTiming().items():
                tm[k] = tm.get(k, 0.0) + v / isteps
        e1.record(stream)
        barrier()
        ms_ing = max_over_ranks(e0.elapsed_time(e1), "cuda") / isteps
        assert ing["status"] == -1 and len(ing["records"]) == ki and ing["dnaLineLen"] == 80
        assert int(ing["records"][:, 3].sum()) == int(sum(lens[:ki]))
        ingest = {"what": "kseq_read_lossless_fasta for a whole multi-FASTA file image on the device (ingest.cu): image in "
                          "pinned host memory -> records + sequences as device texts; H2D of the image inside the timed region",
                  "genomes": ki, "image_bytes": int(img.size), "ms_per_file": ms_ing, "ms": tm,
                  "value": img.size * world / (ms_ing / 1e3) / 1e9, "kernels_value": img.size / (tm["kernels"] / 1e3) / 1e9,
                  "unit": "GB/s", "image": img if rank == 0 else None}

        # the whole hot path from the file image: ingest -> match stage -> stream packing, nothing but the image goes
        # up and nothing but the six streams (and counts) comes back
        def pipeline():
            g = m.ingestFasta(img)
            recs = g["records"]
            m.matchBatchDevice(g["dTexts"], recs[:, 2].copy(), recs[:, 3].copy(), 32, fetch=False)
            return m.packBatch(pd, None, None, loadedPosArr=loaded, lens=recs[:, 3].copy(), raw=True)

        for _ in range(2):
            pipeline()
        barrier()
        e0.record(stream)
        for _ in range(isteps):
            pp_res = pipeline()
        e1.record(stream)
        barrier()
        ms_pipe = max_over_ranks(e0.elapsed_time(e1), "cuda") / isteps
        ingest["pipeline"] = {
            "what": "file image (pinned host memory) -> FASTA ingest -> match stage -> stream packing -> six streams on the "
                    "host: the reference's compression path between file read and back-end coder, all on the device",
            "ms_per_step": ms_pipe, "value": img.size * world / (ms_pipe / 1e3) / 1e9, "unit": "GB/s",
            "h2d_bytes_per_step": int(img.size), "d2h_bytes_per_step": int(sum(pp_res["stream_bytes"].values())),
            "kept_matches": int(pp_res["kept"].sum())}
    ks = max(1, min(16, n))
    lists_sample = m.matchBatchDevice(dbuf.data_ptr(), offs[:ks], lens[:ks], 32)  # copies, for the CPU pack baseline
    state = m.state()
    decode = None
    if not args.no_decode:
        decode = measure_decode(args, m, base, targets, rank)
    out = dict(base=base, targets=targets, limit=limit, sm=sm, n=n, lens=lens, fbytes=fbytes, index_s=index_s, pack=pack,
               index_tm=index_tm, index_warm=index_warm, long_run=long_run, wire=wire, covered_dev=covered_dev, covered_e2e=covered_e2e, nb=nb,
               ingest=ingest, decode=decode,
               lists_sample=lists_sample,
               ms_dev=ms_dev, stage_dev=stage_dev, launches=launches, n_matches=n_matches, counters=counters,
               ms_e2e=ms_e2e, stage_e2e=stage_e2e, table_slots=state["hashSize"], ref_bytes=int(state["pos1"]))
    m.close()
    del dbuf, dev
    torch.cuda.empty_cache()
    return out


def cpu_decode_stage(cfg, full, ref_total, seq_counts, factors, locks, g0_count, n_sample):
    """The reference's own decoder (MBGC_Decoder<false>::decodeReference + decodeTarget, oracle/_ref/libref_decoder.so) on the
    first n_sample targets of the same streams, one host thread (the stock sequential decoder; its parallel mode is the
    lazy one and needs region separators to find independent targets).  Falls back to the oracle port.  -> (seconds, bytes, kind)"""
    ref_so = os.path.join(ROOT, "oracle", "_ref", "libref_decoder.so")
    u8p = C.POINTER(C.c_uint8)
    arrs = [np.frombuffer(x, dtype=np.uint8) if len(x) else np.zeros(8, dtype=np.uint8) for x in full]
    n = len(seq_counts)
    k = max(1, min(n_sample, n))
    pk = cfg.pack
    if os.path.exists(ref_so):
        L = C.CDLL(ref_so)
        L.refd_create.restype = C.c_void_p
        L.refd_create.argtypes = [C.POINTER(C.c_int), C.c_uint64] + [u8p, C.c_uint64] * 6 + [C.POINTER(C.c_uint32), u8p, C.POINTER(C.c_uint64), C.c_uint32]
        L.refd_destroy.argtypes = [C.c_void_p]
        L.refd_decode_reference.argtypes = [C.c_void_p]
        L.refd_decode_target.restype = C.c_uint64
        L.refd_decode_target.argtypes = [C.c_void_p, C.c_uint32, u8p, C.c_uint64]
        pp = (C.c_int * 11)(pk.enableExtensionsWithMismatches, pk.gapDepthOffsetEncoding, pk.gapDepthMismatchesEncoding,
                            pk.lazyDecompressionSupport, pk.frugal64bitLenEncoding, pk.mmsMismatchesInitialScore,
                            pk.mmsMismatchesScoreThreshold, pk.mmsMismatchPenalty, pk.mmsMatchBonus, 1, 1)
        args_ = []
        for a, x in zip(arrs, full):
            args_ += [a.ctypes.data_as(u8p), len(x)]
        counts = (C.c_uint32 * (n + 1))(g0_count, *seq_counts)
        fac = np.array(list(factors) * n, dtype=np.uint8)
        lk = (C.c_uint64 * n)(*locks)
        h = L.refd_create(pp, ref_total, *args_, counts, fac.ctypes.data_as(u8p), lk, n)
        cap = 64 << 20
        out = np.empty(cap, dtype=np.uint8)
        t0 = time.perf_counter()
        L.refd_decode_reference(h)
        total = 0
        for t in range(k):
            total += L.refd_decode_target(h, t, out.ctypes.data_as(u8p), cap)
        dt = time.perf_counter() - t0
        L.refd_destroy(h)
        return dt, int(total), "reference"
    subprocess.run(["make", "-s", "oracle"], cwd=os.path.join(ROOT, "oracle"), check=True)
    L = C.CDLL(os.path.join(ROOT, "oracle", "libmbgc_oracle.so"))

    class PP(C.Structure):
        _fields_ = [(x, C.c_int) for x in ("a", "b", "c", "d", "e", "f", "g", "h", "i", "j", "k", "l", "m", "n")]

    class DS(C.Structure):
        _fields_ = [(f"x{i}", t) for i in range(6) for t in (u8p, C.c_uint64)]
    pp = PP(pk.enableExtensionsWithMismatches, int(pk.gapBreakingMatchMinLength), pk.gapDepthOffsetEncoding, pk.gapDepthMismatchesEncoding,
            pk.lazyDecompressionSupport, pk.frugal64bitLenEncoding, pk.enable40bitReference, pk.mismatchesWithExclusion,
            pk.mmsMismatchesInitialScore, pk.mmsMismatchesScoreThreshold, pk.mmsMismatchPenalty, pk.mmsMatchBonus, 32, factors[0])
    ds = DS()
    for i, (a, x) in enumerate(zip(arrs, full)):
        setattr(ds, f"x{2 * i}", a.ctypes.data_as(u8p))
        setattr(ds, f"x{2 * i + 1}", len(x))
    L.mbo_dec_create.restype = C.c_void_p
    L.mbo_dec_create.argtypes = [C.c_void_p, C.c_uint64, C.c_int, C.c_int, C.c_void_p]
    L.mbo_dec_destroy.argtypes = [C.c_void_p]
    L.mbo_dec_decode_reference.restype = C.c_uint64
    L.mbo_dec_decode_reference.argtypes = [C.c_void_p, C.c_uint32, u8p, C.POINTER(C.c_uint64)]
    L.mbo_dec_decode_target.restype = C.c_uint64
    L.mbo_dec_decode_target.argtypes = [C.c_void_p, C.c_uint32, C.c_int, C.c_int, C.c_uint64, u8p, C.POINTER(C.c_uint64), C.POINTER(C.c_uint64)]
    d = L.mbo_dec_create(C.byref(pp), ref_total, 1, 1, C.byref(ds))
    cap = 64 << 20
    out = np.empty(cap, dtype=np.uint8)
    lens = (C.c_uint64 * 64)()
    unm = (C.c_uint64 * 64)()
    t0 = time.perf_counter()
    L.mbo_dec_decode_reference(d, g0_count, out.ctypes.data_as(u8p), lens)
    total = 0
    for t in range(k):
        total += L.mbo_dec_decode_target(d, seq_counts[t], factors[0], factors[1], locks[t], out.ctypes.data_as(u8p), lens, unm)
    dt = time.perf_counter() - t0
    L.mbo_dec_destroy(d)
    return dt, int(total), "port"


def args_tier_is_near(best):
    return best["batches"] <= 4   # a collection that (almost) never extends the reference


def measure_decode(args, m, base, targets, rank):
    """Row f4: the batch goes through the collection pass (one FASTA image per genome, the reference's sequential schedule)
    and the streams it leaves are decoded again on the device (mbgcg_decode_targets) -- an encode -> decode round trip at the
    bench's full size, every decoded contig compared with the genome it came from.  Runs last: the pass appends separators
    (and, on the divergent tier, the genomes themselves) to the matcher's reference."""
    import torch
    from mbgc_b200.matcher import STREAM_NAMES, SeqConfig

    cfg = SeqConfig.for_mode(args.mode)
    images = [fasta_image([t], rank * len(targets) + 1 + i) for i, t in enumerate(targets)]
    t0 = time.perf_counter()
    streams, info = m.compressSequential(images, cfg)
    enc_s = time.perf_counter() - t0
    del images
    sep = bytes([(ord('"') + 128) & 0xFF])
    full = [base.tobytes() + sep + streams["literals"]] + [streams[k] for k in STREAM_NAMES[1:]]
    ref_total = m.getRefLength()
    factors = (cfg.unmatchedFractionFactor, cfg.unmatchedFractionRCFactor)

    def run():
        t0 = time.perf_counter()
        r = m.decodeTargets(cfg.pack, full, ref_total, info["seqCounts"], factors, g0IsTarget=False, g0SeqCount=1,
                            lockPos=info["lockPos"], fetch=False)
        r["wall_s"] = time.perf_counter() - t0
        return r

    run()
    runs = [run() for _ in range(3)]
    best = min(runs, key=lambda r: r["parse_ms"] + r["copy_ms"])
    dna = int(best["dnaLen"])
    # the decoded contigs back on the host (pinned), compared with what went in
    host = torch.empty(max(dna, 1), dtype=torch.uint8).pin_memory()
    t0 = time.perf_counter()
    m._check(m.lib.mbgcg_decode_fetch(m.h, 0, len(best["lens"]), host.data_ptr()))
    d2h_s = time.perf_counter() - t0
    got = host.numpy()
    at, ok = 0, int(best["lens"][0]) == base.size and bool(np.array_equal(got[:base.size], base))
    at = base.size
    for t, ln in zip(targets, best["lens"][1:]):
        ok = ok and int(ln) == t.size and bool(np.array_equal(got[at:at + t.size], t))
        at += int(ln)
    if not ok:
        raise RuntimeError("decode round trip: the decoded contigs differ from the genomes that were encoded")
    dev_ms = best["parse_ms"] + best["copy_ms"]
    sb = {k: len(v) for k, v in zip(STREAM_NAMES, full)}
    cpu = None
    if rank == 0 and not args.no_cpu:
        ks = 24 if args_tier_is_near(best) else 6
        dt, nb, kind = cpu_decode_stage(cfg, full, ref_total, info["seqCounts"], factors, info["lockPos"], 1, ks)
        cpu = {"value": nb / dt / 1e9, "unit": "GB/s", "cores": 1, "kind": kind,
               "sample": f"decodeReference + decodeTarget for the first {min(ks, len(info['seqCounts']))} targets of the same streams, "
                         f"{nb / 1e6:.0f} MB of output in {dt * 1e3:.0f} ms"}
    return {"cpu_baseline": cpu, "what": "row f4: the six streams of this batch (collection pass, sequential schedule) decoded on the device "
                    "(MBGC_Decoder::decodeTarget for all targets at once); round trip checked byte for byte against the genomes",
            "dna_bytes": dna, "stream_bytes": sb, "contigs": len(best["lens"]), "copy_batches": int(best["batches"]),
            "launches": int(best["launches"]), "parse_ms": best["parse_ms"], "copy_ms": best["copy_ms"],
            "value": dna / (dev_ms / 1e3) / 1e9, "unit": "GB/s of DNA written (CUDA events: stream upload + markers + parse + copy pass)",
            "copy_pass_gbs": dna / (max(best["copy_ms"], 1e-6) / 1e3) / 1e9,
            "copy_pass_hbm_bytes": 2 * dna, "wall_ms": min(r["wall_s"] for r in runs) * 1e3,
            "d2h_ms": d2h_s * 1e3, "e2e_value": dna / (min(r["wall_s"] for r in runs) + d2h_s) / 1e9,
            "encode_pass_s": enc_s, "round_trip_ok": True}


def collection_dir(args):
    return os.path.join(args.scratch, f"mbgc_bench_c3_{args.collection_genomes}_{args.collection_length}")


def measure_collection(args, rank, world, dist, local, peak):
    """C3 of SURVEY.md 8d (BASELINE.json configs[2]): 1000 x 4.9 Mbp near-identical genomes (seed 2025), mode 1, as FILES,
    the fixed collection split over the ranks (strong scaling).  Every rank reads only its own files and runs the
    collection pass of the C ABI on them (mbgcg_seq_*: read -> pinned memory -> device ingest -> match stage -> stream
    packing -> streams on the host, the reference's sequential schedule over speculative windows); the per-rank streams
    concatenated in rank order are the streams of `mbgc -t 1` (tests/test_gpu_multirank.py)."""
    import torch
    import mbgc_b200
    from mbgc_b200.dist import compress_sharded_files, max_over_ranks, shard_range, sum_over_ranks
    from mbgc_b200.matcher import SeqConfig
    from mbgc_b200.synth import _write_one

    n, length, seed = args.collection_genomes, args.collection_length, 2025
    d = collection_dir(args)
    os.makedirs(d, exist_ok=True)
    a, b = shard_range(n, rank, world)
    t0 = time.perf_counter()
    from concurrent.futures import ProcessPoolExecutor
    jobs = [(d, i, length, seed, 1e-4, 1e-5) for i in range(a, b) if not os.path.exists(os.path.join(d, f"g{i:04d}.fna"))]
    if jobs:
        workers = max(1, min(32, (os.cpu_count() or 1) // world))
        with ProcessPoolExecutor(workers) as ex:
            list(ex.map(_write_one, jobs, chunksize=max(1, len(jobs) // (workers * 4))))
    gen_s = time.perf_counter() - t0
    if dist is not None:
        dist.barrier()
    paths = [os.path.join(d, f"g{i:04d}.fna") for i in range(n)]
    shard_bytes = sum(os.path.getsize(p) for p in paths[a:b])
    total_bytes = sum_over_ranks(shard_bytes, "cuda")
    base = synth_base(length, seed)   # = the one contig of g0000.fna, what loadG0Ref puts into the reference (+ RC)
    limit = mgmp_ref_limit(n, max(os.path.getsize(paths[0]), 1 << 21))
    m = mbgc_b200.Matcher(limit, 32, 16, 1, skip_margin_of(1), device=local)
    m.disableSlidingWindow()
    m.loadRef(base, True, True, 0)
    cfg = SeqConfig.for_mode(1)
    pos0 = m.getLoadingPosition()
    readers = int(os.environ.get("MBGC_B200_BENCH_READERS", max(2, min(32, (os.cpu_count() or 1) // world - (3 if world == 1 else 2)))))  # the commit loop, the worker and the stagers need cores too

    def one_pass():
        m.setPosition(pos0, 0)   # back to the state after G0 (a near-identical collection never extends the reference)
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        _, info, ext = compress_sharded_files(m, paths, cfg, device="cuda", readers=readers, keep_streams=False)
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0
        return max_over_ranks(dt, "cuda"), info, ext

    one_pass()                      # warm-up: allocations, pinned arena, page cache
    runs = [one_pass() for _ in range(3)]
    best = min(x[0] for x in runs)
    info = runs[-1][1]
    digest = sum_over_ranks(info["digest"] % (1 << 62), "cuda") % (1 << 62)
    st = info["stats"]
    out = {"what": "BASELINE configs[2] (C3): the fixed 1k-genome collection as FASTA files, sharded over the ranks, every "
                   "rank: files (page cache / tmpfs) -> pinned memory -> FASTA ingest -> match -> pack -> streams on the host",
           "genomes": n, "fasta_bytes": int(total_bytes), "mode": 1, "scaling": "strong", "n_gpus": world,
           "seconds": [x[0] for x in runs], "value": total_bytes / best / 1e9, "unit": "GB/s",
           "timing": "host wall clock between barriers (the region includes file reads and host threads), max over ranks",
           "matches": sum_over_ranks(info["matches"], "cuda"), "unmatched": sum_over_ranks(info["unmatched"], "cuda"),
           "extensions_per_rank": runs[-1][2], "stream_digest": int(digest),
           "rank0": {"files": b - a if rank == 0 else None, "windows": st["windows"], "launches": st["launches"],
                     "ingest_ms": st["ingestMs"], "match_ms": st["matchMs"], "pack_ms": st["packMs"],
                     "wait_staging_ms": st["waitStageMs"], "pageable_bytes": st["pageableBytes"],
                     "stream_bytes": info["stream_bytes"], "reader_threads": readers, "wire_bytes": st.get("wireBytes"),
                     "host": info.get("host")},
           "generate_s": gen_s, "hbm_frac_of_pass": (1.0 * total_bytes / world) / best / 1e9 / peak}
    m.close()
    torch.cuda.empty_cache()
    if not args.no_c4:
        out["c4"] = measure_c4(args, rank, world, dist, local, paths, base, limit, readers, total_bytes)
    return out


def measure_c4(args, rank, world, dist, local, paths, base, limit, readers, total_bytes):
    """C4 of SURVEY.md 8d (BASELINE.json configs[3]): the C3 files in mode 3 -- mode-2 matcher parameters (skip margin 24, RC
    factor 128) plus the reverse-complement redundancy pass over the archive's literal stream
    (SimpleSequenceMatcher::rcMatchSequence, MBGC_Encoder.cpp:637-638).  The collection pass is sharded over the ranks as in C3;
    the literal streams are gathered on rank 0, which runs the RC pass once, as the encoder does."""
    import torch
    import mbgc_b200
    from mbgc_b200.dist import compress_sharded_files, max_over_ranks
    from mbgc_b200.matcher import SeqConfig

    m = mbgc_b200.Matcher(limit, 32, 16, 1, skip_margin_of(3), device=local)
    m.disableSlidingWindow()
    m.loadRef(base, True, True, 0)
    cfg = SeqConfig.for_mode(3)
    pos0 = m.getLoadingPosition()

    def one_pass():
        m.setPosition(pos0, 0)
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        streams, info, ext = compress_sharded_files(m, paths, cfg, device="cuda", readers=readers, keep_streams=True)
        lit = streams["literals"]
        if dist is not None:
            parts = [None] * world
            dist.all_gather_object(parts, lit)
            lit = b"".join(parts)
        rc = None
        if rank == 0:
            sep = bytes([(ord('"') + 128) & 0xFF])
            seq, off, ln, tm = m.rcMatchSequence(np.frombuffer(base.tobytes() + sep + lit, dtype=np.uint8), 55)
            rc = {"literal_bytes_in": len(lit) + base.size + 1, "literal_bytes_out": int(seq.size), "rc_matches": int(off.size // 4),
                  "index_ms": tm["index_ms"], "query_ms": tm["query_ms"]}
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0
        return max_over_ranks(dt, "cuda"), info, rc

    one_pass()
    runs = [one_pass() for _ in range(2)]
    best = min(x[0] for x in runs)
    out = {"what": "BASELINE configs[3] (C4): the C3 files in mode 3 -- collection pass sharded over the ranks, then the "
                   "reverse-complement pass over the gathered literal stream on rank 0 (rcMatchSequence on the device)",
           "mode": 3, "scaling": "strong", "n_gpus": world, "seconds": [x[0] for x in runs], "value": total_bytes / best / 1e9,
           "unit": "GB/s", "rc_pass": runs[-1][2], "matches": runs[-1][1]["matches"]}
    m.close()
    torch.cuda.empty_cache()
    return out


def measure_cli(args, rank, world, dist):
    """Whole-pipeline compression through the reference's own CLI (needs oracle/_ref, built where /root/reference exists):
    stock `mbgc c -m 1` with default threads (the cpu_cli leg of SURVEY.md 8d) beside the batched drop-in on the same
    files, on all `world` GPUs from ONE process (mbgcg_create_multi); rank 0 runs it, the others wait."""
    out = None
    ref_dir = os.path.join(ROOT, "oracle", "_ref")
    have = all(os.path.exists(os.path.join(ref_dir, b)) for b in ("mbgc", "mbgc_b200_batched"))
    if rank == 0 and have:
        sys.path.insert(0, os.path.join(ROOT, "tools"))
        import whole_pipeline
        d = collection_dir(args)
        n = args.collection_genomes
        names = [f"g{i:04d}.fna" for i in range(n)]
        if not args.no_collection and all(os.path.exists(os.path.join(d, x)) for x in names):
            with open(os.path.join(d, "list.txt"), "w") as f:
                f.write("\n".join(names) + "\n")
            total = sum(os.path.getsize(os.path.join(d, x)) for x in names)
            try:
                out = whole_pipeline.measure(d, total, 1, [world], stock_runs=3, skip_t1=False)
                out["genomes"] = n
            except SystemExit as e:
                out = {"error": str(e)}
    elif rank == 0:
        out = {"unavailable": "oracle/_ref/{mbgc,mbgc_b200_batched} not built (needs /root/reference at build time)"}
    if dist is not None:
        dist.barrier()
    return out


def index_build_report(r, peak):
    """B_idx = 1.75 algorithmic bytes per appended reference base (1 read + 0.25 + (4 + 4) / 16 per sample)"""
    tm = r["index_tm"]
    alg = 1.75 * tm["appended"]
    dev_s = max(tm["device_ms"], 1e-6) / 1e3
    warm = {}
    for k, t in (r.get("index_warm") or {}).items():
        a = 1.75 * t["appended"]
        warm[k] = {"device_ms": t["device_ms"], "wall_ms": t["wall_ms"], "achieved": a / max(t["device_ms"], 1e-6) * 1e3 / 1e9,
                   "frac": a / max(t["device_ms"], 1e-6) * 1e3 / 1e9 / peak, "wall_GBps": a / max(t["wall_ms"], 1e-6) * 1e3 / 1e9}
    return {"what": "loadRef of G0 + its reverse complement: staging copy, segment kernels, K2 index inserts.  First call of a "
                    "fresh matcher (pinned staging and mirror pages are allocated inside); `warm` = the same call again on a "
                    "scratch matcher, host mirror on / off",
            "warm": warm,
            "appended_bases": tm["appended"], "alg_bytes": alg, "device_ms": tm["device_ms"], "wall_ms": tm["wall_ms"],
            "roofline": {"bound": "hbm", "kernel": "k_ref_segment + k_index_insert", "achieved": alg / dev_s / 1e9,
                         "peak": peak, "unit": "GB/s", "frac": alg / dev_s / 1e9 / peak,
                         "note": "device time includes the 1 B/base host->device copy of the text (PCIe), which bounds it"},
            "wall_GBps": alg / max(tm["wall_ms"], 1e-6) * 1e3 / 1e9}


def run_b200(args, rank, world):
    import torch

    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dist = None
    if world > 1:
        import torch.distributed as dist_mod
        dist = dist_mod
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))

    sampler = ClockSampler(local)  # rank 0 reports its GPU's clocks; nvidia-smi from every rank would only disturb the timing
    if rank == 0:
        sampler.start()
    r = measure_tier(args, args.tier, rank, world, dist, local)
    sampler.stop_flag = True
    if rank == 0:
        sampler.join(timeout=2)

    steps = args.steps
    n, lens, fbytes, counters = r["n"], r["lens"], r["fbytes"], r["counters"]
    stage_dev = r["stage_dev"]
    # every rank runs batches of the same shape; the timed steps rotate through r["nb"] distinct ones
    value = r["covered_dev"] * world / (r["ms_dev"] / 1e3) / 1e9
    e2e = r["covered_e2e"] * world / (r["ms_e2e"] / 1e3) / 1e9
    peak, peak_src = peaks()
    passes = max(1, counters["passes"])  # every kernel is launched once per pass (at most 1 GiB of text)
    groups = ["anchor", "clean", "fullprobe", "compact", "extend", "resolve"]
    kernels_ms = sum(stage_dev[g] for g in groups)
    with_bytes = [g for g in groups if kernel_alg_bytes(g, counters, r["ref_bytes"])]
    dom = max(with_bytes, key=lambda g: stage_dev[g])  # dominant kernel among those that stream the texts
    dom_ms = stage_dev[dom] / passes
    dom_bytes = kernel_alg_bytes(dom, counters, r["ref_bytes"]) / passes
    achieved = dom_bytes / (dom_ms / 1e3) / 1e9
    stage_achieved = ALG_BYTES_PER_POSITION * counters["positions"] / (kernels_ms / 1e3) / 1e9
    moved = stage_moved_bytes(counters, r["ref_bytes"], r["n_matches"])
    moved_total = sum(moved.values())
    tr = probe_traffic()
    traffic = None
    if tr and STAGE_KERNELS[dom] in tr:
        traffic = tr[STAGE_KERNELS[dom]]["dram_bytes_per_position"] * counters["positions"] / passes
    line = {
        "metric": "compression input GB/s (match stage)", "value": value, "unit": "GB/s", "n_gpus": world,
        "steps": steps, "warmup": args.warmup, "ms_per_step": r["ms_dev"] / steps, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "u8", "data": "synthetic",
        "config": bench_config(args, world),
        "details": {"bases_per_gpu": int(sum(lens)), "fasta_bytes_per_gpu": int(fbytes), "matches_per_gpu": r["n_matches"],
                    "seed_segments": counters["seeds"], "seed_runs": counters["runs"], "chunks": counters["chunks"],
                    "lanes_full_probe": counters["lanes_full_probe"], "ref_limit": int(r["limit"]),
                    "table_slots": r["table_slots"], "skip_margin": r["sm"], "batches_rotated": r["nb"],
                    "timed_region_s": r["ms_dev"] / 1e3},
        "long_run": r["long_run"],
        "roofline": {"bound": "hbm", "kernel": STAGE_KERNELS[dom], "achieved": achieved, "peak": peak, "unit": "GB/s",
                     "frac": achieved / peak, "traffic": traffic, "peak_source": peak_src,
                     "alg_bytes_per_launch": dom_bytes, "launch_ms": dom_ms, "launches_per_step": passes,
                     "share_of_stage": stage_dev[dom] / kernels_ms,
                     "stage_moved": {"what": "all kernels of the match stage on the bytes its own algorithm must move (fast "
                                             "path: ~1.1 B per position + records; see stage_moved_bytes)",
                                     "bytes_per_step": moved_total, "bytes_per_position": moved_total / max(1, counters["positions"]),
                                     "by_group": moved, "achieved": moved_total / (kernels_ms / 1e3) / 1e9,
                                     "frac": moved_total / (kernels_ms / 1e3) / 1e9 / peak, "ms": kernels_ms},
                     "stage_contract": {"what": "the same kernels against the SURVEY 8d contract figure (6.0 B per target "
                                                "position: a probe of every position).  NOT a roofline fraction on the near tier: "
                                                "the fast path never performs those probes",
                                        "contract_GBps": stage_achieved, "ms": kernels_ms},
                     "stage_ms": stage_dev},
        "index_build": index_build_report(r, peak),
        "e2e": {"value": e2e, "unit": "GB/s", "h2d_bytes_per_step": int(sum(lens)),
                "d2h_bytes_per_step": int(r["n_matches"] * 32 + (n + 9) * 4), "ms_per_step": r["ms_e2e"] / steps,
                "wire_bytes_per_step": {"h2d": r["wire"][0], "d2h": r["wire"][1],
                                        "note": "bytes that crossed PCIe: a share of the text travels 2-bit packed (host AVX2 "
                                                "packer), h2d_bytes_per_step is the unpacked size of the inputs"},
                "stage_ms": r["stage_e2e"]},
        "gpu_launches": int(r["launches"]),
        "clocks": sampler.summary(),
    }
    def pack_report(t, world_):
        p = dict(t["pack"])
        p["what"] = ("match stage + MBGC_Encoder::processMatches on the device (K6, pack.cu): texts resident, match lists "
                     "never leave HBM, the six streams copied to the host inside the timed region")
        p["match_plus_pack_value"] = t["fbytes"] * world_ / (p["match_plus_pack_ms_per_step"] / 1e3) / 1e9
        p["unit"] = "GB/s"
        return p

    line["pack"] = pack_report(r, world)
    if r.get("decode"):
        line["decode"] = r["decode"]
    if r["ingest"]:
        image = r["ingest"].pop("image")
        line["ingest"] = r["ingest"]
        # algorithmic bytes of the ingest: every file byte read once, every sequence byte written once
        ing_alg = 2.0 * r["ingest"]["image_bytes"]
        ing_ach = ing_alg / (r["ingest"]["ms"]["kernels"] / 1e3) / 1e9
        line["ingest"]["roofline"] = {"bound": "hbm", "kernel": "k_fa_count + k_fa_starts + k_fa_records + k_fa_check + k_fa_gather",
                                      "achieved": ing_ach, "peak": peak, "unit": "GB/s", "frac": ing_ach / peak,
                                      "alg_bytes_per_launch": ing_alg,
                                      "note": "kernel time includes two host round trips (record count, per-record summaries)"}
        if rank == 0 and world == 1 and not args.no_cpu:
            dt, kind, nrec = cpu_fasta_reader(image)
            line["ingest"]["cpu_baseline"] = {"value": image.size / dt / 1e9, "unit": "GB/s", "cores": 1, "kind": kind,
                                              "sample": f"the same {image.size / 1e6:.0f} MB image, {nrec} records, {dt * 1e3:.0f} ms"}
    if rank == 0 and world == 1 and not args.no_cpu and r["lists_sample"]:
        k = len(r["lists_sample"])
        dt, nb = cpu_pack_stage(r["base"], r["targets"][:k], r["lists_sample"], r["limit"], r["sm"], args.mode)
        sb = sum(fasta_bytes(t.size, HEADER_LEN) for t in r["targets"][:k])
        line["pack"]["cpu_baseline"] = {"value": sb / dt / 1e9, "unit": "GB/s", "cores": 1, "kind": "port",
                                        "sample": f"{k} genomes, mbo_process_matches, {dt * 1e3:.1f} ms, {nb} stream bytes"}
    if rank == 0 and world == 1 and not args.no_cpu:
        sample = r["targets"][:max(1, min(args.cpu_sample, n))]
        sbytes = sum(fasta_bytes(t.size, HEADER_LEN) for t in sample)
        threads = os.cpu_count() or 1
        reps, spent, total, kind, cores = 0, 0.0, 0, "port", 1
        while reps < 3 or (spent < 2.0 and reps < 40):  # a pass is ~10 ms on the near tier: average a few
            dt, kind, total, cores = cpu_match_stage(r["base"], sample, r["limit"], r["sm"], threads)
            spent += dt
            reps += 1
        line["cpu_baseline"] = {"value": sbytes * reps / spent / 1e9, "unit": "GB/s", "cores": cores, "kind": kind,
                                "sample": f"{len(sample)} of the {n} target genomes ({sbytes / 1e6:.0f} MB FASTA), "
                                          f"{reps} passes of matchTexts against the frozen G0+RC index, "
                                          f"{spent / reps * 1e3:.1f} ms each",
                                "matches": total}
    # the collection's other similarity tier (SURVEY.md 8d defines C2 with both), fewer steps
    other = "div" if args.tier == "near" else "near"
    if not args.no_other_tier:
        oargs = argparse.Namespace(**vars(args))
        oargs.steps, oargs.warmup = max(2, args.steps // 2), 3
        o = measure_tier(oargs, other, rank, world, dist, local)
        okern = sum(o["stage_dev"][g] for g in groups)
        line["other_tier"] = {
            "tier": other, "value": o["fbytes"] * world * oargs.steps / (o["ms_dev"] / 1e3) / 1e9,
            "e2e": o["fbytes"] * world * oargs.steps / (o["ms_e2e"] / 1e3) / 1e9, "unit": "GB/s",
            "ms_per_step": o["ms_dev"] / oargs.steps, "matches_per_gpu": o["n_matches"],
            "index_build": index_build_report(o, peak),
            "chunks": o["counters"]["chunks"], "lanes_full_probe": o["counters"]["lanes_full_probe"],
            "stage_frac": ALG_BYTES_PER_POSITION * o["counters"]["positions"] / (okern / 1e3) / 1e9 / peak,
            "stage_ms": o["stage_dev"], "pack": pack_report(o, world)}
        if o.get("decode"):
            line["other_tier"]["decode"] = o["decode"]
        if rank == 0 and world == 1 and not args.no_cpu and o["lists_sample"]:
            k = min(4, len(o["lists_sample"]))
            dt, nb = cpu_pack_stage(o["base"], o["targets"][:k], o["lists_sample"][:k], o["limit"], o["sm"], args.mode)
            sb = sum(fasta_bytes(t.size, HEADER_LEN) for t in o["targets"][:k])
            line["other_tier"]["pack"]["cpu_baseline"] = {
                "value": sb / dt / 1e9, "unit": "GB/s", "cores": 1, "kind": "port",
                "sample": f"{k} genomes, mbo_process_matches, {dt * 1e3:.1f} ms, {nb} stream bytes"}
        if rank == 0 and world == 1 and not args.no_cpu:
            sample = o["targets"][:max(1, min(args.cpu_sample_other, o["n"]))]
            sbytes = sum(fasta_bytes(t.size, HEADER_LEN) for t in sample)
            dt, kind, total, cores = cpu_match_stage(o["base"], sample, o["limit"], o["sm"], os.cpu_count() or 1)
            line["other_tier"]["cpu_baseline"] = {"value": sbytes / dt / 1e9, "unit": "GB/s", "cores": cores,
                                                  "kind": kind, "sample": f"{len(sample)} genomes, one pass, {dt:.2f} s"}
    if not args.no_collection:
        line["collection"] = measure_collection(args, rank, world, dist, local, peak)
    if not args.no_cli:
        cli = measure_cli(args, rank, world, dist)
        if cli is not None:
            line["whole_pipeline"] = cli
    if rank == 0:
        print(json.dumps(line), flush=True)
    if dist is not None:
        dist.destroy_process_group()


def main():
    args = parse_args()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if args.impl == "reference":
        run_reference(args, rank)
    else:
        run_b200(args, rank, world)


if __name__ == "__main__":
    main()
