#!/usr/bin/env python
"""Per-tensor gradient errors of the backward kernels vs autograd through the oracle (debug aid)."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from oracle import render_oracle as orc  # noqa: E402
from sinnerf_b200 import synthetic  # noqa: E402
from sinnerf_b200.nerf import NeRF  # noqa: E402
from sinnerf_b200.rendering import _FieldPass  # noqa: E402

dev = "cuda:0"
p = orc.default_init_params(0)
m = NeRF(use_new_activation=True)
m.load_state_dict(p)
m = m.to(dev)
N = int(os.environ.get("DBG_N", "40")); S = int(os.environ.get("DBG_S", "16"))
rays = synthetic.random_rays("lego", N, seed=1)
z = orc.sample_z(rays[:, 6:7], rays[:, 7:8], S).contiguous()
g = torch.Generator().manual_seed(0)
proj = torch.randn(N, S, 4, generator=g)

po = {k: v.clone().requires_grad_(True) for k, v in p.items()}
dir_enc = orc.embed(rays[:, 3:6], 4)
xyz = (rays[:, None, 0:3] + rays[:, None, 3:6] * z[:, :, None]).reshape(-1, 3)
raw_ref = orc.field_mlp(po, orc.embed(xyz, 10), torch.repeat_interleave(dir_enc, S, dim=0)).view(N, S, 4)
(raw_ref * proj).sum().backward()

raw = _FieldPass.apply(m, rays.to(dev), z.to(dev), *m._param_list())
print("forward rel err", float((raw.detach().cpu() - raw_ref.detach()).norm() / raw_ref.detach().norm()))
(raw * proj.to(dev)).sum().backward()
for k, v in m.named_parameters():
    r = po[k].grad
    e = float((v.grad.cpu() - r).norm() / r.norm().clamp_min(1e-30))
    print(f"{k:32s} |ref| {float(r.norm()):10.4e}  |got| {float(v.grad.norm()):10.4e}  rel err {e:9.3e}")
