"""Seeded synthetic rays of the shapes the reference's datasets emit.

There are no datasets on the build or GPU boxes, so benchmarks and parity tests
use camera-model rays (pixel grid -> pinhole direction -> small random rotation),
not Gaussian noise, so the positional-encoding arguments have realistic size.
Ray rows are ``[ox, oy, oz, dx, dy, dz, near, far]`` and directions are *not*
normalised, as in the reference (datasets/ray_utils.py:73-120).

Shapes:
  lego  400x400, f = 0.5*800/tan(0.5*0.6911)*(400/800), near 2 far 6, white_back
        (datasets/blender_ray_patch_1image_rot3d.py:177,201-211)
  llff  504x378, f = 410, near 1.2, far 7.6, no white_back; patch = 63x84 stride 4
        (datasets/llff_ray_patch_1image_proj.py:351,396-403,625-630)
  dtu   640x512, d = [(i-cx)/fx, (j-cy)/fy, 1], near 2.125 far 4.525, white_back
        (datasets/dtu_proj.py:17-34,290,312,396-398)
"""
from __future__ import annotations

import math
from dataclasses import dataclass
from typing import Dict

import torch


@dataclass(frozen=True)
class FrameShape:
    name: str
    width: int
    height: int
    focal: float
    near: float
    far: float
    white_back: bool
    opencv: bool  # True: +z forward, y down (DTU); False: -z forward, y up (blender/LLFF)


SHAPES = {
    "lego": FrameShape("lego", 400, 400, 0.5 * 800 / math.tan(0.5 * 0.6911) * (400 / 800), 2.0, 6.0, True, False),
    "llff": FrameShape("llff", 504, 378, 410.0, 1.2, 7.6, False, False),
    "dtu": FrameShape("dtu", 640, 512, 2892.33 * (640 / 1600), 2.125, 4.525, True, True),
}


def _small_rotation(gen: torch.Generator) -> torch.Tensor:
    """Rodrigues rotation about a random axis by <= ~11 degrees."""
    axis = torch.randn(3, generator=gen, dtype=torch.float64)
    axis = axis / axis.norm()
    ang = (torch.rand((), generator=gen, dtype=torch.float64) - 0.5) * 0.4
    kx, ky, kz = axis.tolist()
    K = torch.tensor([[0, -kz, ky], [kz, 0, -kx], [-ky, kx, 0]], dtype=torch.float64)
    R = torch.eye(3, dtype=torch.float64) + math.sin(ang) * K + (1 - math.cos(ang)) * (K @ K)
    return R.to(torch.float32)


def frame_rays(shape: str, seed: int = 0, rows=None, cols=None) -> torch.Tensor:
    """(H*W, 8) fp32 rays of a whole frame (or of the given row/col index sets), row-major
    over (row, col) like the reference's flattened (H, W) grids."""
    fs = SHAPES[shape]
    gen = torch.Generator().manual_seed(seed)
    R = _small_rotation(gen)
    origin = torch.tensor([0.0, 0.0, 4.0 if not fs.opencv else -3.3]) + (torch.rand(3, generator=gen) - 0.5) * 0.2
    jj = torch.arange(fs.height, dtype=torch.float32) if rows is None else rows.to(torch.float32)
    ii = torch.arange(fs.width, dtype=torch.float32) if cols is None else cols.to(torch.float32)
    j, i = torch.meshgrid(jj, ii, indexing="ij")
    if fs.opencv:
        d = torch.stack([(i - fs.width / 2) / fs.focal, (j - fs.height / 2) / fs.focal, torch.ones_like(i)], -1)
    else:
        d = torch.stack([(i - fs.width / 2) / fs.focal, -(j - fs.height / 2) / fs.focal, -torch.ones_like(i)], -1)
    d = d.reshape(-1, 3) @ R.t()
    n = d.shape[0]
    rays = torch.empty(n, 8, dtype=torch.float32)
    rays[:, 0:3] = origin + (torch.rand(n, 3, generator=gen) - 0.5) * 0.0  # one camera centre per frame
    rays[:, 3:6] = d
    rays[:, 6] = fs.near
    rays[:, 7] = fs.far
    return rays


def patch_rays(shape: str, patch_h: int, patch_w: int, stride: int, seed: int = 0) -> torch.Tensor:
    """Strided patch of a frame, as the *_ray_patch_* datasets cut it
    (datasets/llff_ray_patch_1image_proj.py:625-646)."""
    fs = SHAPES[shape]
    gen = torch.Generator().manual_seed(seed + 7919)
    span_h, span_w = (patch_h - 1) * stride + 1, (patch_w - 1) * stride + 1
    top = int(torch.randint(0, fs.height - span_h + 1, (1,), generator=gen))
    left = int(torch.randint(0, fs.width - span_w + 1, (1,), generator=gen))
    rows = torch.arange(top, top + span_h, stride)
    cols = torch.arange(left, left + span_w, stride)
    return frame_rays(shape, seed, rows, cols)


def random_rays(shape: str, n: int, seed: int = 0) -> torch.Tensor:
    """n rays at random pixels of a frame (the 4096-ray random sets of a training step)."""
    fs = SHAPES[shape]
    allr = frame_rays(shape, seed)
    gen = torch.Generator().manual_seed(seed + 104729)
    idx = torch.randint(0, fs.width * fs.height, (n,), generator=gen)
    return allr[idx].contiguous()


def default_init_params(seed: int) -> Dict[str, torch.Tensor]:
    """State dict of `torch.manual_seed(seed); NeRF(use_new_activation=True)` -- the seeded default-init weights the bench
    and the measurement tools render with (the modules are created in the reference's order, models/nerf.py:66-103, so
    these are the reference's numbers; tests/test_oracle_golden.py holds them equal to the oracle's own copy).  The global
    RNG state is left untouched."""
    from .nerf import NeRF
    state = torch.random.get_rng_state()
    torch.manual_seed(seed)
    sd = {k: v.detach().clone() for k, v in NeRF(use_new_activation=True).state_dict().items()}
    torch.random.set_rng_state(state)
    return sd
