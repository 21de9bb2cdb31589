#!/usr/bin/env python
"""render_rays (inference, HBM-resident rays) on the shapes BASELINE.json names besides the bench line:
configs[2] LLFF 63x84 stride-4 patch in bf16, configs[3] one 8-GPU shard of the 640x512 DTU frame.

    python tools/time_configs.py
"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from sinnerf_b200.synthetic import default_init_params  # noqa: E402  (seeded default-init weights)
from sinnerf_b200 import synthetic  # noqa: E402
from sinnerf_b200.nerf import NeRF, Embedding  # noqa: E402
from sinnerf_b200.rendering import render_rays  # noqa: E402

dev = torch.device("cuda:0")
models = []
for seed in (0, 1):
    m = NeRF(use_new_activation=True)
    m.load_state_dict(default_init_params(seed))
    models.append(m.to(dev))
emb = [Embedding(3, 10), Embedding(3, 4)]


def bench(name, rays, white_back, precision, iters=20):
    rays = rays.to(dev)
    with torch.no_grad():
        for _ in range(3):
            render_rays(models, emb, rays, 64, False, 0, 0, 64, 32768, white_back, precision=precision)
        torch.cuda.synchronize()
        ts = []
        for _ in range(iters):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            render_rays(models, emb, rays, 64, False, 0, 0, 64, 32768, white_back, precision=precision)
            e1.record()
            torch.cuda.synchronize()
            ts.append(e0.elapsed_time(e1))
    ms = sorted(ts)[len(ts) // 2]
    print(f"{name:58s} {precision:6s} rays={rays.shape[0]:7d}  {ms:8.3f} ms  {rays.shape[0] / ms / 1e3:7.3f} M rays/s")


patch = synthetic.patch_rays("llff", 63, 84, 4, seed=0)
shard = synthetic.frame_rays("dtu", seed=0)[:40960]
for prec in ("bf16", "f16x3"):
    bench("configs[2] LLFF 63x84 stride-4 patch", patch, False, prec)
for prec in ("f16x3", "bf16"):
    bench("configs[3] DTU 640x512 frame, one of 8 contiguous shards", shard, True, prec)
