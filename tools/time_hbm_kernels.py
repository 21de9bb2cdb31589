#!/usr/bin/env python
"""Time the HBM-bound stage kernels alone (CUDA events, L2 flushed between runs) and print achieved
GB/s = algorithmic bytes / time against the measured copy bandwidth (MEASURED_PEAKS.json).

    python tools/time_hbm_kernels.py [--rays 160000]
"""
import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

from sinnerf_b200 import _lib, synthetic  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--rays", type=int, default=160000)
args = ap.parse_args()
dev = torch.device("cuda:0")
lib = _lib.load()
peak = 6650.0
try:
    peak = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"]
except Exception:
    pass
n = args.rays
rays = synthetic.frame_rays("lego", seed=0)[:n].to(dev)
st = _lib.stream_ptr(dev)
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)


def timed(fn, reps=5):
    fn()
    torch.cuda.synchronize()
    best = 1e9
    for _ in range(reps):
        flush.fill_(1)
        if not os.environ.get("SNB_DIRTY_FLUSH"):
            # read the buffer back: L2 then holds CLEAN lines of it.  Writing alone leaves ~126 MB of dirty lines whose
            # write-back lands inside the timed kernel (a ~100 us kernel then measures ~25 % slow)
            flush.sum()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        fn()
        e1.record()
        torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1))
    return best


def report(name, ms, nbytes):
    gbs = nbytes / ms / 1e6
    print(f"{name:28s} {ms * 1e3:9.1f} us  {nbytes / 1e6:9.1f} MB algorithmic  {gbs:8.1f} GB/s  {gbs / peak:5.2f} of measured HBM copy peak ({peak:.0f} GB/s)")


for S in (64, 128):
    P = n * S
    z = (torch.linspace(2, 6, S, device=dev)[None, :] + torch.rand(n, 1, device=dev) * 0.01).contiguous()
    raw = torch.rand(n, S, 4, device=dev)
    raw[..., 3] = torch.randn(n, S, device=dev) * 3
    rgb, depth, w = torch.empty(n, 3, device=dev), torch.empty(n, device=dev), torch.empty(n, S, device=dev)
    ms = timed(lambda: lib.snb_composite_forward(_lib.ptr(raw), 4, _lib.ptr(z), _lib.ptr(rays), None, 0.0, 1, n, S,
                                                 _lib.ptr(rgb), _lib.ptr(depth), _lib.ptr(w), st))
    report(f"composite_fwd S={S}", ms, P * 24 + n * 48)
    g_rgb, g_d, g_raw = torch.randn(n, 3, device=dev), torch.randn(n, device=dev), torch.empty(n, S, 4, device=dev)
    ms = timed(lambda: lib.snb_composite_backward(_lib.ptr(raw), _lib.ptr(z), _lib.ptr(rays), None, 0.0, 1, _lib.ptr(g_rgb),
                                                  _lib.ptr(g_d), None, n, S, _lib.ptr(g_raw), st))
    report(f"composite_bwd S={S}", ms, P * 36 + n * 48)

S, Ni = 64, 64
z = (torch.linspace(2, 6, S, device=dev)[None, :] + torch.rand(n, 1, device=dev) * 0.01).contiguous()
w = torch.rand(n, S, device=dev) ** 4
u = torch.linspace(0, 1, Ni, device=dev)
zf = torch.empty(n, S + Ni, device=dev)
ms = timed(lambda: lib.snb_importance_merge(_lib.ptr(z), _lib.ptr(w), _lib.ptr(u), 0, n, S, Ni, 1e-5, _lib.ptr(zf), None, st))
report("importance_merge 64+64", ms, n * (8 * S + 4 * (S + Ni)))
steps = torch.linspace(0, 1, S, device=dev)
zc = torch.empty(n, S, device=dev)
ms = timed(lambda: lib.snb_sample_coarse(_lib.ptr(rays), _lib.ptr(steps), None, 0.0, 0, n, S, _lib.ptr(zc), st))
report("sample_coarse S=64", ms, n * (32 + 4 * S))
P = n * 128
x = (torch.rand(P, 3, device=dev) - 0.5) * 6
out = torch.empty(P, 63, device=dev)
ms = timed(lambda: lib.snb_embed(_lib.ptr(x), P, 3, 10, _lib.ptr(out), st))
report("embed (3,10) P=20.5M", ms, P * (12 + 252))
