#!/usr/bin/env python
"""Render one BASELINE shape N times (no timing, no profiler): the command ncu wraps for a launch list
(`ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file launches.csv python tools/render_once.py ...`;
tools/launch_summary.py aggregates the second half = the last render when --n 2).

    python tools/render_once.py [--shape frame|patch|dtu] [--precision f16x3] [--n 2]
"""
import argparse
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from sinnerf_b200 import rendering, synthetic  # noqa: E402
from sinnerf_b200.nerf import Embedding, NeRF  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--shape", default="frame")
ap.add_argument("--precision", default="f16x3")
ap.add_argument("--n", type=int, default=2)
args = ap.parse_args()
dev = torch.device("cuda:0")
models = []
for i in (0, 1):
    m = NeRF(use_new_activation=True)
    m.load_state_dict(synthetic.default_init_params(i))
    models.append(m.to(dev))
emb = [Embedding(3, 10), Embedding(3, 4)]
rays = {"frame": lambda: synthetic.frame_rays("lego", seed=0), "dtu": lambda: synthetic.frame_rays("dtu", seed=0),
        "patch": lambda: synthetic.patch_rays("llff", 63, 84, 4, seed=0)}[args.shape]().to(dev)
for _ in range(args.n):
    with torch.no_grad():
        out = rendering.render_rays(models, emb, rays, 64, False, 0, 0, 64, 32768, args.shape != "patch", precision=args.precision)
    torch.cuda.synchronize()
print(f"{args.shape} {args.precision}: {rays.shape[0]} rays, rgb_fine mean {float(out['rgb_fine'].mean()):.6f}")
