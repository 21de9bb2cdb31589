#!/usr/bin/env python
"""Training keeps the reference's scheme (Lightning DDP -> NCCL gradient all-reduce, reference train.py:51-52):
wrap the two NeRF modules in DistributedDataParallel, let every rank render its own ray set through
render_rays (tensor-core training path) and check that the all-reduced gradients equal the mean of the
per-rank gradients computed without DDP.

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 tools/ddp_check.py
"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402
from torch.nn.parallel import DistributedDataParallel as DDP  # noqa: E402

from sinnerf_b200.synthetic import default_init_params  # noqa: E402  (seeded default-init weights)
from sinnerf_b200 import synthetic  # noqa: E402
from sinnerf_b200.nerf import NeRF, Embedding  # noqa: E402
from sinnerf_b200.rendering import render_rays  # noqa: E402

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
emb = [Embedding(3, 10), Embedding(3, 4)]


def fresh():
    ms = []
    for seed in (0, 1):
        m = NeRF(use_new_activation=True)
        m.load_state_dict(default_init_params(seed))
        ms.append(m.to(dev))
    return ms


def loss_for(models, r):
    rays = synthetic.random_rays("lego", 512, seed=40 + r).to(dev)
    target = torch.rand(512, 3, generator=torch.Generator().manual_seed(7 + r)).to(dev)
    out = render_rays(models, emb, rays, 64, False, 0, 0, 64, 32768, True)
    return ((out["rgb_fine"] - target) ** 2).mean() + ((out["rgb_coarse"] - target) ** 2).mean() + 0.1 * out["depth_fine"].mean()


class Step(torch.nn.Module):
    """Stand-in for the LightningModule (reference models/sinnerf.py): owns both NeRFs, forward = one loss."""

    def __init__(self, ms):
        super().__init__()
        self.nerf_coarse, self.nerf_fine = ms

    def forward(self, r):
        return loss_for([self.nerf_coarse, self.nerf_fine], r)


# DDP: every rank its own rays; the wrapped module's forward arms the reducer, render_rays reads the same leaves
models = fresh()
ddp = DDP(Step(models), device_ids=[local])
ddp(rank).backward()
ddp_grads = [p.grad.detach().clone() for m in models for p in m.parameters()]

# reference: the same ray sets on this rank alone, averaged
ref_models = fresh()
acc = None
for r in range(world):
    for m in ref_models:
        m.zero_grad(set_to_none=True)
    loss_for(ref_models, r).backward()
    g = [p.grad.detach().clone() for m in ref_models for p in m.parameters()]
    acc = g if acc is None else [a + b for a, b in zip(acc, g)]
ref_grads = [a / world for a in acc]
worst = 0.0
for a, b in zip(ddp_grads, ref_grads):
    if float(b.norm()) == 0.0:
        continue
    worst = max(worst, float((a - b).norm() / b.norm()))
ok = worst <= 1e-4
print(f"rank {rank}: DDP all-reduced gradients vs mean of per-rank gradients: worst rel-L2 {worst:.2e} -> {'ok' if ok else 'MISMATCH'}", flush=True)
dist.barrier()
dist.destroy_process_group()
sys.exit(0 if ok else 1)
