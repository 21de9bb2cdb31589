"""Host-side logic of the multi-GPU path on CPU: world_size-2 gloo processes shard a frame,
render their slab (with the ORACLE standing in for the CUDA renderer -- tests may do that, the
product never does) and all-gather the pixels; the result must equal the single-process render."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from sinnerf_b200.distributed import pack_pixels, render_rays_sharded, shard_bounds


def test_shard_bounds_cover_and_are_disjoint():
    for n in (0, 1, 7, 8, 9, 160000, 327680, 5292):
        for world in (1, 2, 3, 4, 8):
            spans = [shard_bounds(n, world, r) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            for (a, b), (c, d) in zip(spans, spans[1:]):
                assert b == c and a <= b and c <= d
            assert max(b - a for a, b in spans) == (-(-n // world) if n else 0)


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, n_rays, tmp):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    torch.set_num_threads(2)
    from oracle import render_oracle as orc
    from sinnerf_b200 import synthetic
    rays = synthetic.random_rays("dtu", n_rays, seed=4)
    pc, pf = orc.default_init_params(0), orc.default_init_params(1)

    def render(r):
        with torch.no_grad():
            return orc.render_rays(pc, pf, r, N_samples=16, N_importance=8, noise_std=0.0, white_back=True)

    out = render_rays_sharded(render, rays)
    torch.save(out, os.path.join(tmp, f"out{rank}.pt"))
    dist.destroy_process_group()


@pytest.mark.parametrize("n_rays", [37, 64])
def test_sharded_render_equals_single_process(tmp_path, n_rays):
    world = 2
    mp.spawn(_worker, args=(world, _free_port(), n_rays, str(tmp_path)), nprocs=world, join=True)
    from oracle import render_oracle as orc
    from sinnerf_b200 import synthetic
    rays = synthetic.random_rays("dtu", n_rays, seed=4)
    with torch.no_grad():
        ref = pack_pixels(orc.render_rays(orc.default_init_params(0), orc.default_init_params(1), rays,
                                          N_samples=16, N_importance=8, noise_std=0.0, white_back=True))
    outs = [torch.load(os.path.join(tmp_path, f"out{r}.pt")) for r in range(world)]
    assert torch.equal(outs[0], outs[1])
    assert outs[0].shape == (n_rays, 4)
    # slabs are rendered independently; per-ray results do not depend on the batch they ride in
    assert torch.allclose(outs[0], ref, rtol=0, atol=1e-6)


def _gather_worker(rank, world, port, tmp):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from sinnerf_b200.distributed import PixelGather
    pg = PixelGather(5, torch.device("cpu"))
    seen = []
    outs = []
    for frame in range(5):          # more frames than buffers: slots are reused only after their gather finished
        local = torch.full((5, 4), float(10 * frame + rank))
        outs.append((frame, pg.submit(local)))
        if len(outs) == 2:           # consume the older result before its slot comes up again
            f, buf = outs.pop(0)
            pg.wait_all()
            seen.append((f, buf.clone()))
    pg.wait_all()
    seen += [(f, b.clone()) for f, b in outs]
    torch.save(seen, os.path.join(tmp, f"g{rank}.pt"))
    dist.destroy_process_group()


def test_pixel_gather_double_buffering(tmp_path):
    """PixelGather (the asynchronous, double-buffered all-gather behind the weak-scaling bench): every frame's
    gathered buffer holds all ranks' slabs of THAT frame, in rank order, on every rank."""
    world = 2
    mp.spawn(_gather_worker, args=(world, _free_port(), str(tmp_path)), nprocs=world, join=True)
    for r in range(world):
        seen = torch.load(os.path.join(tmp_path, f"g{r}.pt"))
        assert [f for f, _ in seen] == list(range(5))
        for f, buf in seen:
            assert buf.shape == (10, 4)
            assert torch.equal(buf[:5], torch.full((5, 4), float(10 * f))) and torch.equal(buf[5:], torch.full((5, 4), float(10 * f + 1)))


class _FakePeerPixels:
    """The protocol of distributed.PeerPixels with ONE shared CPU buffer per frame slot standing in for the symmetric
    memory of all ranks (the stores that the CUDA kernel makes are done by the fake render function below)."""

    def __init__(self, rows, world, rank, shared):
        self.rows, self.world, self.rank, self.bufs = rows, world, rank, shared
        self.k, self.log = 0, []

    def begin(self):
        self.k += 1
        self.log.append(("begin", self.k - 1))
        return self.k - 1

    def scatter(self, k, row_offset):
        return [("slot", k % len(self.bufs))], row_offset

    def commit(self, k):
        self.log.append(("commit", k))

    def frame(self, k):
        self.log.append(("frame", k))
        return self.bufs[k % len(self.bufs)]


def test_render_frame_p2p_places_every_slab_and_keeps_the_call_order():
    """render_frame_p2p = begin -> render own slab with (destinations, row offset of the slab) -> commit -> frame; the
    ranks' slabs tile the frame exactly (ragged tail included) and rows past the frame are never written."""
    from sinnerf_b200.distributed import render_frame_p2p
    for n, world in ((10, 3), (9, 4), (5, 8), (64, 2)):
        shared = [torch.full((n + 3, 4), -1.0) for _ in range(4)]
        rays = torch.arange(n * 8, dtype=torch.float32).reshape(n, 8)
        pps = [_FakePeerPixels(n + 3, world, r, shared) for r in range(world)]
        for frame_no in range(3):
            outs = []
            for r in range(world):
                def render(slab, sc, r=r):
                    (dst,), off = sc
                    assert dst == ("slot", frame_no % 4)
                    shared[dst[1]][off:off + slab.shape[0]] = slab[:, :4] + 100.0 * frame_no     # what the kernel's stores do
                outs.append(render_frame_p2p(render, rays, pps[r]))
            for o in outs:
                assert o.shape == (n, 4) and torch.equal(o, rays[:, :4] + 100.0 * frame_no)
            assert bool((shared[frame_no % 4][n:] == -1.0).all())
        for pp in pps:
            assert pp.log == [(w, k) for k in range(3) for w in ("begin", "commit", "frame")]
    with pytest.raises(ValueError):
        render_frame_p2p(lambda slab, sc: None, torch.zeros(20, 8), _FakePeerPixels(10, 2, 0, [torch.zeros(10, 4)] * 4))
