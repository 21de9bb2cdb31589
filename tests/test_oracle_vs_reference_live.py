"""Randomised differential test: the CPU oracle against the REFERENCE ITSELF, imported from /root/reference.

The committed goldens (tests/golden/, tests/test_oracle_golden.py) are what travels; this file widens the
pin where the reference is importable -- the build container -- and is skipped everywhere else (the GPU box
has no /root/reference; nothing marked `gpu` may touch it).  Seeds, sizes and option combinations beyond
the goldens: use_disp, white_back, perturb/noise (the reference draws from the global generator in the
order rand, randn, rand, randn -- the oracle must consume it identically), N_importance = 0, test_time.
"""
import os
import sys

import pytest
import torch

REF = "/root/reference"
pytestmark = pytest.mark.skipif(not os.path.isdir(os.path.join(REF, "models")), reason="reference tree not present")

from oracle import render_oracle as orc  # noqa: E402
from sinnerf_b200 import synthetic  # noqa: E402


@pytest.fixture(scope="module")
def ref():
    sys.path.insert(0, REF)
    try:
        from models.nerf import NeRF, Embedding
        from models.rendering import render_rays, sample_pdf
    finally:
        sys.path.remove(REF)
    return {"NeRF": NeRF, "Embedding": Embedding, "render_rays": render_rays, "sample_pdf": sample_pdf}


def ref_models(ref, params):
    out = []
    for p in params:
        m = ref["NeRF"](use_new_activation=True)
        m.load_state_dict(p)
        out.append(m.eval())
    return out


CASES = [
    # shape, n, S, Ni, use_disp, perturb, noise_std, white_back, seed
    ("lego", 33, 64, 64, False, 0.0, 0.0, True, 1),
    ("llff", 20, 48, 24, False, 1.0, 1.0, False, 2),
    ("dtu", 17, 32, 16, True, 1.0, 0.0, True, 3),
    ("lego", 9, 64, 0, False, 0.0, 1.0, False, 4),
    ("llff", 5, 16, 40, True, 0.0, 0.0, False, 5),
]


@pytest.mark.parametrize("shape,n,S,Ni,use_disp,perturb,noise_std,white_back,seed", CASES)
def test_render_rays_oracle_equals_live_reference(ref, shape, n, S, Ni, use_disp, perturb, noise_std, white_back, seed):
    rays = synthetic.random_rays(shape, n, seed=seed)
    pc, pf = orc.default_init_params(10 + seed), orc.default_init_params(20 + seed)
    models = ref_models(ref, [pc, pf])
    emb = [ref["Embedding"](3, 10), ref["Embedding"](3, 4)]
    with torch.no_grad():
        torch.manual_seed(100 + seed)
        want = ref["render_rays"](models, emb, rays, S, use_disp, perturb, noise_std, Ni, 1024, white_back, test_time=False)
        torch.manual_seed(100 + seed)
        got = orc.render_rays(pc, pf if Ni > 0 else None, rays, N_samples=S, N_importance=Ni, use_disp=use_disp, perturb=perturb,
                              noise_std=noise_std, white_back=white_back)
    for k, v in want.items():
        assert k in got, k
        assert got[k].shape == v.shape, k
        err = float((got[k] - v).abs().max())
        scale = max(float(v.abs().max()), 1e-6)
        assert err <= 2e-5 * scale, (k, err, scale)


def test_test_time_keys_and_values(ref):
    rays = synthetic.random_rays("lego", 12, seed=9)
    pc, pf = orc.default_init_params(1), orc.default_init_params(2)
    models = ref_models(ref, [pc, pf])
    emb = [ref["Embedding"](3, 10), ref["Embedding"](3, 4)]
    with torch.no_grad():
        want = ref["render_rays"](models, emb, rays, 64, False, 0, 0, 64, 1024, True, test_time=True)
        got = orc.render_rays(pc, pf, rays, N_samples=64, N_importance=64, noise_std=0.0, white_back=True, test_time=True)
    assert set(k for k in got if not k.startswith("_")) == set(want)
    for k, v in want.items():
        assert float((got[k] - v).abs().max()) <= 2e-5 * max(float(v.abs().max()), 1e-6), k


@pytest.mark.parametrize("seed", [0, 1, 2])
def test_sample_pdf_oracle_equals_live_reference(ref, seed):
    g = torch.Generator().manual_seed(seed)
    n, m, ni = 19, 23 + seed, 31
    bins = torch.sort(torch.rand(n, m + 1, generator=g) * 4 + 2, dim=-1).values
    w = torch.rand(n, m, generator=g) ** 3
    w[0] = 0.0                                    # all-zero weights row (the eps path)
    want = ref["sample_pdf"](bins, w, ni, det=True)
    got = orc.sample_pdf(bins, w, ni, det=True)
    # identical arithmetic; allow the inverse-CDF's knot discontinuity (SURVEY hard part 3) on a few samples
    diff = (got - want).abs()
    assert float(diff.median()) == 0.0
    assert int((diff > 1e-5).sum()) <= 4


def test_autograd_oracle_equals_live_reference(ref):
    """Gradients of a random projection of all outputs w.r.t. all 48 parameter tensors: reference autograd vs
    autograd through the oracle (perturb and noise on: the sample_pdf detach and the RNG order both matter)."""
    rays = synthetic.random_rays("llff", 14, seed=21)
    pc, pf = orc.default_init_params(31), orc.default_init_params(32)
    models = ref_models(ref, [pc, pf])
    for m in models:
        m.train()
    emb = [ref["Embedding"](3, 10), ref["Embedding"](3, 4)]
    torch.manual_seed(77)
    want = ref["render_rays"](models, emb, rays, 32, False, 1.0, 1.0, 24, 1024, False, test_time=False)
    g = torch.Generator().manual_seed(5)
    proj = {k: torch.randn(v.shape, generator=g) for k, v in want.items()}
    sum((want[k] * proj[k]).sum() for k in want).backward()
    oc = {k: v.clone().requires_grad_(True) for k, v in pc.items()}
    of = {k: v.clone().requires_grad_(True) for k, v in pf.items()}
    torch.manual_seed(77)
    got = orc.render_rays(oc, of, rays, N_samples=32, N_importance=24, perturb=1.0, noise_std=1.0, white_back=False)
    sum((got[k] * proj[k]).sum() for k in want).backward()
    for params, model in ((oc, models[0]), (of, models[1])):
        sd = dict(model.named_parameters())
        for k, v in params.items():
            a, b = v.grad, sd[k].grad
            assert (a is None) == (b is None) or float(a.abs().sum()) == 0.0 or float(b.abs().sum()) == 0.0, k
            if a is None or b is None:
                continue
            assert float((a - b).norm()) <= 2e-4 * max(float(b.norm()), 1e-12), (k, float((a - b).norm() / b.norm()))
