#!/usr/bin/env python
"""Kernel timeline (CUPTI via torch.profiler) of one render_rays call: start / duration of every kernel and the idle gaps
between them -- where a short render (the configs[2] patch, < 1 ms) spends the time that is not the field kernels.

    python tools/profile_render.py [--precision bf16] [--shape patch|frame]
"""
import argparse
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402
from torch.profiler import ProfilerActivity, profile  # noqa: E402

from sinnerf_b200.synthetic import default_init_params  # noqa: E402  (seeded default-init weights)
from sinnerf_b200 import rendering, synthetic  # noqa: E402
from sinnerf_b200.nerf import Embedding, NeRF  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--precision", default="bf16")
ap.add_argument("--shape", default="patch")
args = ap.parse_args()
dev = torch.device("cuda:0")
models = []
for i in (0, 1):
    m = NeRF(use_new_activation=True)
    m.load_state_dict(default_init_params(i))
    models.append(m.to(dev))
emb = [Embedding(3, 10), Embedding(3, 4)]
rays = (synthetic.patch_rays("llff", 63, 84, 4, seed=0) if args.shape == "patch" else synthetic.frame_rays("lego", seed=0)).to(dev)
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)


def step():
    with torch.no_grad():
        return rendering.render_rays(models, emb, rays, 64, False, 0, 0, 64, 32768, False, precision=args.precision)


for _ in range(3):
    step()
torch.cuda.synchronize()
with profile(activities=[ProfilerActivity.CPU, ProfilerActivity.CUDA]) as prof:
    for _ in range(2):
        flush.fill_(1)
        torch.cuda.synchronize()
        step()
        torch.cuda.synchronize()
evs = sorted((e for e in prof.events() if e.device_type == torch.autograd.DeviceType.CUDA), key=lambda e: e.time_range.start)
# the second step: kernels after the last flush fill
last_fill = max(i for i, e in enumerate(evs) if "fill" in e.name.lower() or "memset" in e.name.lower() or "FillFunctor" in e.name)
evs = evs[last_fill + 1:]
t0 = evs[0].time_range.start
prev_end = t0
print(f"{'start us':>9s} {'dur us':>8s} {'gap us':>7s}  kernel")
tot_k = 0.0
for e in evs:
    s, d = e.time_range.start - t0, e.time_range.end - e.time_range.start
    print(f"{s:9.1f} {d:8.1f} {e.time_range.start - prev_end:7.1f}  {e.name[:90]}")
    prev_end = e.time_range.end
    tot_k += d
print(f"span {prev_end - t0:.1f} us, kernels {tot_k:.1f} us, gaps {prev_end - t0 - tot_k:.1f} us")
cpu = [e for e in prof.events() if e.device_type == torch.autograd.DeviceType.CPU and e.name.startswith("cudaLaunchKernel")]
print(f"{len(cpu)} cudaLaunchKernel calls in the two profiled steps")
