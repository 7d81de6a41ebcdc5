"""GPU: host texts that travel 2-bit packed (hostpack.cpp + k_unpack_blocks / k_patch_exceptions) must arrive on
the device byte for byte, and the match lists must not depend on how the text travelled."""
import numpy as np
import pytest

from scenarios import CudaFace
from util import OracleMatcher, mgmp_ref_limit, synth_collection

pytestmark = pytest.mark.gpu


def _texts():
    gs = synth_collection(7, 2_000_000, 321, 2e-4, 2e-5)
    ts = [g.copy() for g in gs[1:]]
    ts[0][100_000:130_000] = ord("N")                       # long N run
    ts[1][500_001:500_020] = np.frombuffer(b"RYKMSWBDHVNrykmswbd", dtype=np.uint8)
    ts[2][1_000_000:1_300_000] = np.frombuffer(bytes(ts[2][1_000_000:1_300_000]).lower(), dtype=np.uint8)  # raw block
    ts[3] = ts[3][:1_234_567]                                # odd length
    ts.append(ts[4][:5])
    ts.append(np.zeros(0, np.uint8))
    return gs[0], ts


@pytest.mark.parametrize("pinned", [False, True], ids=["pageable", "pinned"])
@pytest.mark.parametrize("raw_permille", ["0", "250", "1000"])
def test_packed_transfer_is_exact(monkeypatch, pinned, raw_permille):
    import torch
    monkeypatch.setenv("MBGC_B200_PACK", "1")
    monkeypatch.setenv("MBGC_B200_RAW_PERMILLE", raw_permille)
    monkeypatch.setenv("MBGC_B200_PACK_THREADS", "5")
    base, ts = _texts()
    keep = []
    if pinned:
        for i, t in enumerate(ts):
            if t.size:
                p = torch.from_numpy(t.copy()).pin_memory()
                keep.append(p)
                ts[i] = p.numpy()
    limit = mgmp_ref_limit(8, base.size)
    g = CudaFace(limit)
    o = OracleMatcher(limit)
    try:
        for m in (g, o):
            m.disable_sliding_window()
            m.load_ref(base, rc=True)
        got = g.match_batch(ts)
        for i, t in enumerate(ts):
            assert np.array_equal(g.m.stagedText(i, t.size), t), f"text {i} did not arrive intact"
        for a, t in zip(got, ts):
            assert np.array_equal(a, o.match(t))
        monkeypatch.setenv("MBGC_B200_PACK", "0")
        plain = g.match_batch(ts)
        for a, b in zip(got, plain):
            assert np.array_equal(a, b)
    finally:
        g.close()
        o.close()
