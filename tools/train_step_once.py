#!/usr/bin/env python
"""The training step bench.py times as extra.configs_4_ddp (render_rays_multi forward + backward, fused per-ray losses,
FusedAdam + weight re-pack), single GPU, 3 warm-up + 5 timed steps: the command ncu wraps for the step's launch list.

    ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file launches.csv python tools/train_step_once.py
"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

import bench  # noqa: E402
from sinnerf_b200.nerf import Embedding, NeRF  # noqa: E402
from sinnerf_b200.synthetic import default_init_params  # noqa: E402

dev = torch.device("cuda:0")
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
res = bench.bench_configs_4_train(dev, 0, 0, 1, "f16x3", flush, torch.cuda.synchronize, NeRF, Embedding, default_init_params)
print({k: res[k] for k in ("ms_per_step", "peak_mem_gib", "iters")})
