#!/usr/bin/env python
"""Time the NeRF part of a SinNeRF training step (BASELINE.json configs[4] shape): four render_rays
calls of 4096 rays each (64+64 samples, perturb=1, noise_std=1) forward + backward, MSE-style loss.

    python tools/time_train.py [--rays 4096] [--calls 4] [--iters 3]
"""
import argparse
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from sinnerf_b200.synthetic import default_init_params  # noqa: E402  (seeded default-init weights)
from sinnerf_b200 import synthetic  # noqa: E402
from sinnerf_b200.nerf import NeRF, Embedding  # noqa: E402
from sinnerf_b200.rendering import render_rays, render_rays_multi  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--rays", type=int, default=4096)
ap.add_argument("--calls", type=int, default=4)
ap.add_argument("--iters", type=int, default=3)
ap.add_argument("--multi", action="store_true", help="one render_rays_multi call instead of --calls separate ones")
args = ap.parse_args()
dev = torch.device("cuda:0")
models = []
for seed in (0, 1):
    m = NeRF(use_new_activation=True)
    m.load_state_dict(default_init_params(seed))
    models.append(m.to(dev))
emb = [Embedding(3, 10), Embedding(3, 4)]
batches = [synthetic.random_rays("lego", args.rays, seed=i).to(dev) for i in range(args.calls)]
target = torch.rand(args.rays, 3, device=dev)


def step():
    for m in models:
        m.zero_grad(set_to_none=True)
    loss = 0.0
    outs = render_rays_multi(models, emb, batches, 64, False, 1.0, 1.0, 64, 32768, True) if args.multi else \
        [render_rays(models, emb, r, 64, False, 1.0, 1.0, 64, 32768, True) for r in batches]
    for out in outs:
        loss = loss + ((out["rgb_coarse"] - target) ** 2).mean() + ((out["rgb_fine"] - target) ** 2).mean() \
            + 0.1 * out["depth_fine"].mean()
    loss.backward()
    return loss


step()
torch.cuda.synchronize()
ts = []
for _ in range(args.iters):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    step()
    e1.record()
    torch.cuda.synchronize()
    ts.append(e0.elapsed_time(e1))
ms = min(ts)
n = args.rays * args.calls
flops = 3 * 2 * 593408 * n * 192
print(f"train step: {args.calls} x {args.rays} rays fwd+bwd  ms={ms:.1f}  {n / ms * 1e3:.0f} rays/s  "
      f"{flops / ms / 1e9:.1f} TFLOP/s (3x forward FLOPs)  peak mem {torch.cuda.max_memory_allocated() / 2**30:.1f} GiB")
