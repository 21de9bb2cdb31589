#!/usr/bin/env python
"""torchrun check of the pixel-scatter path (>= 2 GPUs): frames assembled by the compositing kernel's stores into peer /
multicast memory (distributed.PeerPixels) against the NCCL all-gather of the same slabs (render_rays_sharded), bitwise,
over several back-to-back frames (buffer reuse), then the time per 327 680-ray frame of the three exchange schemes.

    python -m torch.distributed.run --nproc-per-node 2 --master-addr 127.0.0.1 tools/p2p_check.py
"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402

from sinnerf_b200.synthetic import default_init_params  # noqa: E402  (seeded default-init weights)
from sinnerf_b200 import synthetic  # noqa: E402
from sinnerf_b200.distributed import PeerPixels, PixelGather, pack_pixels, render_frame_p2p, render_rays_sharded, shard_bounds  # noqa: E402
from sinnerf_b200.nerf import Embedding, NeRF  # noqa: E402
from sinnerf_b200.rendering import render_rays  # noqa: E402

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
models = []
for i in (0, 1):
    m = NeRF(use_new_activation=True)
    m.load_state_dict(default_init_params(i))
    models.append(m.to(dev))
emb = [Embedding(3, 10), Embedding(3, 4)]
frame = synthetic.frame_rays("dtu", seed=0).to(dev)


def render(r, sc=None):
    with torch.no_grad():
        return render_rays(models, emb, r, 64, False, 0, 0, 64, 32768, True, pixel_scatter=sc)


def log(msg):
    if rank == 0:
        print(msg, flush=True)


# ---- correctness: 7 frames of 20 001 rays (odd: ragged slabs), different rays per frame, both addressing modes
n = 20001
ok = True
for mc in (True, False):
    pp = PeerPixels(n, dev, multicast=mc)
    log(f"PeerPixels: world {world}, multicast {'on' if pp.multicast else 'off'}")
    got = []
    for f in range(7):
        rays = frame[f * 1000:f * 1000 + n]
        got.append(render_frame_p2p(render, rays, pp).clone())     # the read is enqueued before begin() of frame f + 2
    for f in range(7):
        rays = frame[f * 1000:f * 1000 + n]
        want = render_rays_sharded(render, rays)
        same = bool(torch.equal(got[f], want))
        ok &= same
        if not same:
            print(f"[{rank}] frame {f} multicast={mc}: MISMATCH max|d| {float((got[f] - want).abs().max()):.3e}", flush=True)
    torch.cuda.synchronize()
    dist.barrier()
flag = torch.tensor([1 if ok else 0], device=dev)
dist.all_reduce(flag, op=dist.ReduceOp.MIN)
log(f"pixel scatter == NCCL all-gather, bitwise, 7 pipelined frames x 2 addressing modes, all ranks: {bool(flag.item())}")

# ---- timing: the 327 680-ray frame, strong-scaled, 12 frames back to back
n = frame.shape[0]
lo, hi = shard_bounds(n, world, rank)
per = -(-n // world)


def timed(fn, finish, iters=12):
    for _ in range(2):
        fn()
    finish()
    torch.cuda.synchronize()
    dist.barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        fn()
    finish()
    e1.record()
    torch.cuda.synchronize()
    t = torch.tensor([e0.elapsed_time(e1) / iters], device=dev)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t)


ms_block = timed(lambda: render_rays_sharded(render, frame), lambda: None)
pg = PixelGather(per, dev)


def gather_step():
    local_px = pack_pixels(render(frame[lo:hi]))
    if local_px.shape[0] < per:
        local_px = torch.cat([local_px, local_px.new_zeros((per - local_px.shape[0], 4))], dim=0)
    pg.submit(local_px)


ms_async = timed(gather_step, pg.wait_all)
res = {}
for mc in (True, False):
    pp = PeerPixels(n, dev, multicast=mc)
    res[mc] = timed(lambda: render_frame_p2p(render, frame, pp), pp.wait_all)
log(f"327 680-ray frame over {world} GPUs, ms per frame (max over ranks): blocking all-gather {ms_block:.3f}, "
    f"async all-gather (PixelGather) {ms_async:.3f}, kernel stores to the multicast address {res[True]:.3f}, "
    f"kernel stores to each peer {res[False]:.3f}")
dist.barrier()
dist.destroy_process_group()
