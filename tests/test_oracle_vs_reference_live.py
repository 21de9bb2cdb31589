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
