"""GPU parity of the on-device ray generation against the reference-made golden and the oracle."""
import numpy as np
import pytest
import torch

from oracle import render_oracle as orc
from tests._common import load_npz

pytestmark = pytest.mark.gpu


def t(x):
    return torch.from_numpy(np.asarray(x).copy())


def test_camera_rays_vs_reference_golden():
    from sinnerf_b200.rays import camera_rays
    gz = load_npz("rays.npz")
    for tag in ("lego", "llff"):
        H, W, f, near, far = gz[f"{tag}_cfg"]
        got = camera_rays(int(H), int(W), float(f), t(gz[f"{tag}_c2w"]), near, far).cpu()
        ref = t(gz[f"{tag}_rays"])
        assert got.shape == ref.shape
        assert torch.equal(got[:, [0, 1, 2, 6, 7]], ref[:, [0, 1, 2, 6, 7]])          # origin, near, far: exact
        assert (got[:, 3:6] - ref[:, 3:6]).abs().max() <= 1e-6                         # 3-term dot products
    H, W, fx, fy, cx, cy, near, far = gz["dtu_cfg"]
    got = camera_rays(int(H), int(W), (float(fx), float(fy)), t(gz["dtu_c2w"]), near, far,
                      center=(float(cx), float(cy)), opencv=True).cpu()
    assert (got - t(gz["dtu_rays"])).abs().max() <= 1e-6


def test_camera_rays_strided_window_and_errors():
    from sinnerf_b200.rays import camera_rays
    c2w = torch.tensor([[1., 0., 0., 0.1], [0., 1., 0., -0.2], [0., 0., 1., 4.0]])
    win = (10, 20, 63, 84, 4)          # the 63x84 stride-4 LLFF patch
    got = camera_rays(378, 504, 410.0, c2w, 1.2, 7.6, window=win).cpu()
    ref = orc.camera_rays(378, 504, 410.0, c2w, 1.2, 7.6, window=win)
    assert got.shape == (63 * 84, 8) and (got - ref).abs().max() <= 1e-6
    full = camera_rays(378, 504, 410.0, c2w, 1.2, 7.6).cpu().view(378, 504, 8)
    assert torch.equal(got.view(63, 84, 8), full[10:10 + 63 * 4:4, 20:20 + 84 * 4:4])   # window == strided slice
    with pytest.raises(ValueError):
        camera_rays(378, 504, 410.0, c2w, 1.2, 7.6, window=(300, 0, 63, 84, 4))
    with pytest.raises(RuntimeError):
        camera_rays(4, 4, 1.0, c2w, 1.0, 2.0, device="cpu")


def test_render_camera_equals_render_rays_on_uploaded_rays():
    from sinnerf_b200.nerf import NeRF, Embedding
    from sinnerf_b200.rays import camera_rays, render_camera
    from sinnerf_b200.rendering import render_rays
    models = []
    for seed in (0, 1):
        m = NeRF(use_new_activation=True)
        m.load_state_dict(orc.default_init_params(seed))
        models.append(m.to("cuda:0"))
    emb = [Embedding(3, 10), Embedding(3, 4)]
    c2w = torch.tensor([[1., 0., 0., 0.], [0., 1., 0., 0.], [0., 0., 1., 4.0]])
    with torch.no_grad():
        a = render_camera(models, emb, 24, 32, 40.0, c2w, 2.0, 6.0, N_samples=64, N_importance=64, noise_std=0,
                          white_back=True)
        rays = orc.camera_rays(24, 32, 40.0, c2w, 2.0, 6.0).to("cuda:0")
        b = render_rays(models, emb, rays, 64, False, 0, 0, 64, 32768, True)
    assert torch.allclose(a["rgb_fine"], b["rgb_fine"], atol=2e-5) and a["rgb_fine"].shape == (24 * 32, 3)
    assert torch.equal(camera_rays(24, 32, 40.0, c2w, 2.0, 6.0)[:, :3], rays[:, :3])
