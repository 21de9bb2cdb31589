"""CPU tests: the oracle (oracle/render_oracle.py) against fixtures produced by the
reference itself (tests/golden/make_golden.py) and the known-answer vectors of
SURVEY.md section 8c.  These pin the oracle; the GPU tests then pin the CUDA path to it."""
import numpy as np
import pytest
import torch

from oracle import render_oracle as orc
from tests._common import (RENDER_CASES, assert_close, case_cfg, case_params, case_rng, load_npz,
                           room_params)

ST = load_npz("stages.npz")


def t(x):
    return torch.from_numpy(np.asarray(x).copy())


def test_embed_matches_reference_bitwise():
    x = t(ST["embed_x"])
    assert torch.equal(orc.embed(x, 10), t(ST["embed_xyz_out"]))
    assert torch.equal(orc.embed(x, 4), t(ST["embed_dir_out"]))


def test_embed_known_answer():
    x = torch.tensor([[0.1, 0.2, 0.3]])
    e = orc.embed(x, 2)
    expect = torch.cat([x, torch.sin(x), torch.cos(x), torch.sin(2 * x), torch.cos(2 * x)], -1)
    assert torch.equal(e, expect)
    assert torch.equal(e, t(ST["embed_L2_kat"]))


def test_activations_match_reference():
    x = t(ST["act_x"])
    assert torch.equal(orc.shifted_softplus(x), t(ST["act_softplus"]))
    assert torch.equal(orc.widened_sigmoid(x), t(ST["act_wsigmoid"]))
    assert float(orc.widened_sigmoid(torch.tensor(-100.0))) == pytest.approx(-0.001, abs=1e-7)
    assert float(orc.widened_sigmoid(torch.tensor(100.0))) == pytest.approx(1.001, abs=1e-7)


@pytest.mark.parametrize("tag,params", [("seed0", None), ("room_coarse", "coarse"), ("room_fine", "fine")])
def test_field_mlp_matches_reference(tag, params):
    p = orc.default_init_params(0) if params is None else room_params(params)
    x = t(ST["mlp_in"])
    out = orc.field_mlp(p, x[:, :63], x[:, 63:])
    assert_close(out, ST[f"mlp_{tag}_out"], 2e-6, tag)
    if tag == "seed0":
        s = orc.field_mlp(p, x[:, :63], None, sigma_only=True)
        assert_close(s, ST["mlp_seed0_sigma"], 2e-6, "sigma_only")
        assert torch.equal(s[:, 0], out[:, 3])


def test_sample_pdf_known_answers():
    bins = torch.tensor([[0., 1., 2., 3., 4.]])
    kat = {"ones": ([1., 1., 1., 1.], 5, [0, 1, 2, 3, 4]),
           "spike": ([0., 0., 1., 0.], 5, [0, 2.2499876, 2.4999950, 2.7500024, 4]),
           "zero": ([0., 0., 0., 0.], 5, [0, 1, 2, 3, 4]),
           "ramp": ([.1, .2, .3, .4], 8, [0, 1.2142537, 1.9285324, 2.4285479, 2.9047413, 3.2857037,
                                          3.6428518, 4])}
    for tag, (w, n, expect) in kat.items():
        got = orc.sample_pdf(bins, torch.tensor([w]), n, det=True)
        assert torch.allclose(got, torch.tensor([expect], dtype=torch.float32), atol=2e-6), tag
        assert torch.equal(got, t(ST[f"pdf_kat_{tag}"])), tag


def test_sample_pdf_matches_reference_bitwise():
    bins, w = t(ST["pdf_bins"]), t(ST["pdf_w"])
    assert torch.equal(orc.sample_pdf(bins, w, 64, det=True), t(ST["pdf_det_out"]))
    assert torch.equal(orc.sample_pdf(bins, w, 64, det=False, u=t(ST["pdf_rand_u"])), t(ST["pdf_rand_out"]))


def test_composite_known_answer():
    # constant field sigma=0.5, c=0.25, ray d=(0,0,-2), near 2 far 6, S=4, white_back
    z = orc.sample_z(torch.tensor([[2.0]]), torch.tensor([[6.0]]), 4)
    sigma = torch.full((1, 4), 0.5)
    rgb = torch.full((1, 4, 3), 0.25)
    rgb_map, depth, w = orc.composite(sigma, z, torch.tensor([[2.0]]), rgb, None, True)
    assert torch.allclose(w, torch.tensor([[0.7364029, 0.1941137, 0.0511678, 0.0183156]]), atol=1e-6)
    assert float(depth) == pytest.approx(2.4685283, abs=1e-5)
    assert torch.allclose(rgb_map, torch.full((1, 3), 0.2500001), atol=1e-6)


@pytest.mark.parametrize("name", RENDER_CASES)
def test_render_rays_matches_reference(name):
    case = load_npz(f"render_{name}.npz")
    cfg = case_cfg(case)
    coarse, fine = case_params(case)
    with torch.no_grad():
        out = orc.render_rays(coarse, fine, t(case["rays"]), rng=case_rng(case), **cfg)
    keys = [k[4:] for k in case if k.startswith("out_")]
    assert sorted(keys) == sorted(out.keys())
    for k in keys:
        # same ATen kernels in the same order on the same CPU -> agreement to fp32 round-off;
        # (not bitwise: the oracle's point chunking differs from the reference's)
        assert_close(out[k], case["out_" + k], 5e-6, f"{name}:{k}")


def test_rng_draw_order_matches_reference():
    """With rng=None the oracle draws from torch's generator in the reference's order."""
    case = load_npz("render_llff_room_64p64_train.npz")
    coarse, fine = case_params(case)
    torch.manual_seed(1234)           # the seed make_golden.py used
    with torch.no_grad():
        out = orc.render_rays(coarse, fine, t(case["rays"]), **case_cfg(case))
    assert_close(out["rgb_fine"], case["out_rgb_fine"], 5e-6, "rgb_fine")


def test_properties_weights():
    case = load_npz("render_llff_room_64p64.npz")
    w = t(case["out_opacity_fine"])
    assert (w >= 0).all() and (w.sum(-1) <= 1 + 1e-5).all()


def test_oracle_autograd_matches_reference_autograd():
    """Gradients: autograd through the oracle == autograd through the reference (golden made by
    tests/golden/make_golden.py::grad_golden), including the detach of the importance samples."""
    gz = load_npz("grad_llff_room_train.npz")
    pc = {k: v.clone().requires_grad_(True) for k, v in room_params("coarse").items()}
    pf = {k: v.clone().requires_grad_(True) for k, v in room_params("fine").items()}
    rng = {k[4:]: t(v) for k, v in gz.items() if k.startswith("rng_")}
    out = orc.render_rays(pc, pf, t(gz["rays"]), N_samples=64, N_importance=64, perturb=1.0, noise_std=1.0, rng=rng)
    loss = sum((out[k[5:]] * t(gz[k])).sum() for k in sorted(gz) if k.startswith("proj_"))
    assert float(loss) == pytest.approx(float(gz["loss"]), rel=1e-5)
    loss.backward()
    for which, params in (("coarse", pc), ("fine", pf)):
        for name, prm in params.items():
            ref_norm = float(gz[f"gnorm_{which}/{name}"])
            assert float(prm.grad.norm()) == pytest.approx(ref_norm, rel=2e-4, abs=1e-9), (which, name)
            key = f"grad_{which}/{name}"
            if key in gz:
                assert_close(prm.grad, gz[key], 2e-4, key)


def test_camera_rays_match_reference():
    """oracle ray generation == the reference's get_ray_directions/get_rays (golden rays.npz)."""
    gz = load_npz("rays.npz")
    for tag in ("lego", "llff"):
        H, W, f, near, far = gz[f"{tag}_cfg"]
        got = orc.camera_rays(int(H), int(W), float(f), t(gz[f"{tag}_c2w"]), near, far)
        assert torch.allclose(got, t(gz[f"{tag}_rays"]), rtol=0, atol=1e-6), tag
    H, W, fx, fy, cx, cy, near, far = gz["dtu_cfg"]
    got = orc.camera_rays(int(H), int(W), (float(fx), float(fy)), t(gz["dtu_c2w"]), near, far, center=(float(cx), float(cy)),
                          opencv=True)
    assert torch.allclose(got, t(gz["dtu_rays"]), rtol=0, atol=1e-6)


def test_synthetic_default_init_equals_the_oracles():
    """bench.py and tools/ take their seeded weights from sinnerf_b200.synthetic (the product side must not import the
    oracle); they are the same numbers as the oracle's restatement of the reference's default init."""
    import torch
    from oracle import render_oracle as orc
    from sinnerf_b200 import synthetic
    before = torch.random.get_rng_state()
    for seed in (0, 1, 5):
        a, b = synthetic.default_init_params(seed), orc.default_init_params(seed)
        assert list(a) == list(b)
        assert all(torch.equal(a[k], b[k]) for k in a)
    assert torch.equal(before, torch.random.get_rng_state())
