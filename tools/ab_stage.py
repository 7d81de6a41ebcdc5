"""A/B of kernel variants on the bench's match stage (device-resident batches): MBGC_B200_* switches are read at every
launch, so one process times them all on the same inputs and checks that the match lists do not change.
    python tools/ab_stage.py [--tier near|div] [--steps 40]"""
import argparse, hashlib, json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import mbgc_b200
from mbgc_b200.synth import synth_base, synth_mutant, mgmp_ref_limit

ap = argparse.ArgumentParser()
ap.add_argument("--tier", default="near")
ap.add_argument("--steps", type=int, default=40)
ap.add_argument("--genomes", type=int, default=100)
ap.add_argument("--e2e", action="store_true", help="time mbgcg_match_batch with HOST buffers (pinned) instead of the device-resident call")
ap.add_argument("--configs", default="CLEAN_VARIANT=0,SCAN_LOOKBACK=1;CLEAN_VARIANT=1,SCAN_LOOKBACK=1;CLEAN_VARIANT=2,SCAN_LOOKBACK=1;CLEAN_VARIANT=0,SCAN_LOOKBACK=0;CLEAN_VARIANT=1,SCAN_LOOKBACK=0")
args = ap.parse_args()
TIERS = {"near": (1e-4, 1e-5), "div": (1e-2, 1e-3)}
SEED = 2025
p_snp, p_indel = TIERS[args.tier]
base = synth_base(5_000_000, SEED)
nb = 2 if args.tier == "near" else 1
batches = [[synth_mutant(base, SEED * 1000003 + i, p_snp, p_indel) for i in range(1 + b * args.genomes, 1 + (b + 1) * args.genomes)] for b in range(nb)]
m = mbgc_b200.Matcher(mgmp_ref_limit(args.genomes + 1, base.size), 32, 16, 1, 16, device=0)
m.disableSlidingWindow()
m.loadRef(base, True, True, 0)
dev = []
host = [[torch.from_numpy(t).pin_memory() for t in tg] for tg in batches] if args.e2e else None
for tg in batches:
    offs, cur = [], 64
    for t in tg:
        offs.append(cur)
        cur = (cur + t.size + 64 + 15) // 16 * 16
    dbuf = torch.zeros(cur + 8192, dtype=torch.uint8, device="cuda")
    for t, o in zip(tg, offs):
        dbuf[o:o + t.size].copy_(torch.from_numpy(t))
    dev.append((dbuf, np.array(offs, dtype=np.uint64), np.array([t.size for t in tg], dtype=np.uint64)))
torch.cuda.synchronize()
stream = torch.cuda.ExternalStream(m.stream())

def step(k):
    if args.e2e:
        return m.matchBatch([h.numpy() for h in host[k % nb]], 32, copy=False, split=False)
    d, o, l = dev[k % nb]
    return m.matchBatchDevice(d.data_ptr(), o, l, 32, copy=False, split=False)

ref_digest = None
for cfg in args.configs.split(";"):
    for kv in cfg.split(","):
        k, v = kv.split("=")
        os.environ["MBGC_B200_" + k] = v
    m.setStageTiming(False)
    for k in range(6):
        step(k)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    for k in range(args.steps):
        res = step(k)
    e1.record(stream)
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / args.steps
    m.setStageTiming(True)
    stage = {}
    for k in range(10):
        step(k)
        for name, v in m.timing().items():
            stage[name] = stage.get(name, 0.0) + v / 10
    res = step(0)
    dig = hashlib.md5(np.ascontiguousarray(res[0]).tobytes()).hexdigest()
    if ref_digest is None:
        ref_digest = dig
    extra = {"GBps": round(sum(t.size + 70 + t.size // 80 for t in batches[0]) / ms / 1e6, 1), "wire": m.wireBytes()} if args.e2e else {}
    print(json.dumps({**extra, "config": cfg, "ms_per_step": round(ms, 4), "matches": int(len(res[0])), "same_lists": dig == ref_digest,
                      "stage_ms": {k: round(v, 4) for k, v in stage.items()}}), flush=True)
