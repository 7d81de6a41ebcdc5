"""Where the wall time of the batched drop-in goes outside the reference's own timers (MBGC_B200_TRACE stamps), on C3."""
import os, re, subprocess, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from mbgc_b200.synth import write_collection
d = "/dev/shm/c3ab"
names, total = write_collection(d, 1000, 4_900_000, 2025, 1e-4, 1e-5)
open(os.path.join(d, "list.txt"), "w").write("\n".join(names) + "\n")
R = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "oracle", "_ref")
import torch
torch.zeros(1, device="cuda")   # keep the GPU initialised, as bench.py does while it runs the CLI
for rep in range(6):
    e = dict(os.environ); e["MBGC_B200_TRACE"] = "1"; e["MBGC_B200_FAST_EXIT"] = str(rep % 2)
    time.sleep(1.0)   # let the previous process's teardown finish in the driver
    t0 = time.time()
    r = subprocess.run([os.path.join(R, "mbgc_b200_batched"), "c", "-Q", "-m", "1", "-f", "list.txt", d + "/o.mbgc"], cwd=d, capture_output=True, text=True, env=e)
    t1 = time.time()
    log = r.stdout + r.stderr
    st = {m.group(1).strip(): float(m.group(2)) for m in re.finditer(r"stamp (.+?)\s+(\d+\.\d+)", log)}
    enc = re.search(r"encoding time - (\d+)", log)
    ctx = re.search(r"create\s+\d+ bytes\s+([0-9.]+) ms", log)
    print("fast_exit", rep % 2, "wall %.3f enc %s create %s | start->shim %.3f | shim->atexit %.3f | atexit->destroy %.3f | destroy %.3f | destroy->exit %.3f" % (
        t1 - t0, enc.group(1) if enc else None, ctx.group(1) if ctx else None, st.get("shim initialised", t0) - t0,
        st.get("main returned (on_exit)", 0) - st.get("shim initialised", 0), st.get("destroy begins", 0) - st.get("main returned (on_exit)", 0),
        st.get("destroy ends", 0) - st.get("destroy begins", 0), t1 - st.get("destroy ends", t1)), flush=True)
    print("   ", sorted(st.items(), key=lambda kv: kv[1]))
