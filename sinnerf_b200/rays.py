"""On-device ray generation: the step right before `render_rays`.

Replaces the reference's CPU pipeline `get_ray_directions` -> `get_rays` -> `torch.cat([o, d, near, far])`
(datasets/ray_utils.py:73-120; DTU variant datasets/dtu_proj.py:17-34) and the `.cuda()` upload of the
(H*W, 8) tensor (eval.py:155) with one kernel that writes the rays where the renderer reads them.
"""
from __future__ import annotations

import ctypes as C
from typing import Optional, Sequence, Tuple

import torch

from . import _lib


def camera_rays(H: int, W: int, focal, c2w, near: float, far: float, *, center: Optional[Sequence[float]] = None,
                opencv: bool = False, window: Optional[Tuple[int, int, int, int, int]] = None,
                device="cuda") -> torch.Tensor:
    """(N,8) rays `[o, d, near, far]` of a pinhole camera, row-major over (row, col).

    focal: float or (fx, fy); c2w: (3,4) camera-to-world; center: (cx, cy), default (W/2, H/2);
    opencv=False -> d = [(i-cx)/fx, -(j-cy)/fy, -1] (blender / LLFF, ray_utils.py:87-89),
    opencv=True  -> d = [(i-cx)/fx, (j-cy)/fy, 1] (DTU, dtu_proj.py:31-32);
    window = (row0, col0, rows, cols, stride): the strided patch the *_ray_patch_* datasets cut
    (llff_ray_patch_1image_proj.py:625-646); default = the whole frame.
    """
    dev = torch.device(device)
    if dev.type != "cuda":
        raise RuntimeError("sinnerf_b200.rays.camera_rays: expected a CUDA device (this package has no CPU path)")
    fx, fy = (float(focal), float(focal)) if not isinstance(focal, (tuple, list)) else (float(focal[0]), float(focal[1]))
    cx, cy = (W / 2, H / 2) if center is None else (float(center[0]), float(center[1]))
    row0, col0, rows, cols, stride = (0, 0, H, W, 1) if window is None else window
    if row0 + (rows - 1) * stride >= H or col0 + (cols - 1) * stride >= W:
        raise ValueError("camera_rays: window leaves the frame")
    m = torch.as_tensor(c2w, dtype=torch.float32).reshape(3, 4).contiguous().cpu()
    arr = (C.c_float * 12)(*m.flatten().tolist())
    rays = torch.empty(rows * cols, 8, device=dev, dtype=torch.float32)
    _lib.require_device(rays, "camera_rays")
    with torch.cuda.device(dev):
        _lib.check(_lib.load().snb_generate_rays(arr, fx, fy, cx, cy, float(near), float(far), int(opencv), row0, col0,
                                                 rows, cols, stride, _lib.ptr(rays), _lib.stream_ptr(dev)),
                   "snb_generate_rays")
    return rays


def render_camera(models, embeddings, H, W, focal, c2w, near, far, *, center=None, opencv=False, window=None,
                  device="cuda", **render_kwargs):
    """`render_rays` on the rays of a camera generated on the device (no (H*W,8) host tensor, no H2D copy)."""
    from .rendering import render_rays
    rays = camera_rays(H, W, focal, c2w, near, far, center=center, opencv=opencv, window=window, device=device)
    return render_rays(models, embeddings, rays, **render_kwargs)
