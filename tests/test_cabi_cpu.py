"""CPU tests of the C-ABI boundary: the library loads, exports every symbol the header declares,
and validates arguments without touching a GPU.  No compute is launched here."""
import ctypes as C
import os
import re

import pytest

from sinnerf_b200 import _lib, build

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def lib():
    build.build()
    return _lib.load()


def header_symbols():
    src = open(os.path.join(ROOT, "include", "sinnerf_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(snb_[a-z_0-9]+)\s*\(", src)))


def test_header_and_binding_agree():
    assert header_symbols() == sorted(_lib.SIGNATURES)


def test_every_declared_symbol_is_exported(lib):
    for name in header_symbols():
        assert hasattr(lib, name), name


def test_version(lib):
    assert lib.snb_version() == 100


def test_packed_sizes(lib):
    # fp32 image: header + padded K-major weights + biases + heads
    n_w = 64 * 256 + 3 * 256 * 256 + 320 * 256 + 3 * 256 * 256 + 256 * 256 + 288 * 128
    n_b = 9 * 256 + 128
    assert lib.snb_packed_weights_bytes(0) == 256 + 4 * (n_w + n_b + 256 + 4 + 384 + 4)
    assert lib.snb_packed_weights_bytes(99) == 0


def test_argument_validation_without_gpu(lib):
    rc = lib.snb_composite_forward(None, 3, None, None, None, 0.0, 0, 4, 64, None, None, None, None)
    assert rc == -1
    assert b"raw_channels" in lib.snb_last_error()
    rc = lib.snb_sample_coarse(None, None, None, 1.0, 0, 4, 64, None, None)
    assert rc == -1
    rc = lib.snb_importance_merge(None, None, None, 0, 4, 2, 8, 1e-5, None, None, None)
    assert rc == -1 and b"N_samples >= 3" in lib.snb_last_error()
    assert lib.snb_render_forward(None, None) == -1
    rc = lib.snb_composite_forward_scatter(None, None, None, None, 0.0, 0, 0, 64, None, None, None, None, None)
    assert rc == -1 and b"scatter" in lib.snb_last_error()
    sc = _lib.SnbPixelScatter()
    sc.n_dst = _lib.MAX_PIXEL_DST + 1
    rc = lib.snb_composite_forward_scatter(None, None, None, None, 0.0, 0, 0, 64, None, None, None, C.byref(sc), None)
    assert rc == -1 and b"destinations" in lib.snb_last_error()
    a = _lib.SnbRenderArgs()
    a.n_rays = 0
    a.n_samples = 64
    assert lib.snb_render_forward(C.byref(a), None) == 0     # empty input is a no-op
    # training entry points: precision / null checks come before any CUDA call; empty passes are no-ops
    assert lib.snb_field_forward_train(None, 99, None, None, 4, 64, None, None, None, None, None, None) == -1
    assert b"precision" in lib.snb_last_error()
    assert lib.snb_field_forward_train(None, 1, None, None, 4, 64, None, None, None, None, None, None) == -1
    assert b"null pointer" in lib.snb_last_error()
    assert lib.snb_field_forward_train(None, 1, None, None, 0, 64, None, None, None, None, None, None) == 0
    nul = (C.c_void_p * 24)()
    assert lib.snb_field_backward(nul, nul, 1, None, None, None, None, None, None, 0, None, None, None, None, None, None) == -1
    assert b"is null" in lib.snb_last_error()
    assert lib.snb_field_backward(None, None, 1, None, None, None, None, None, None, 0, None, None, None, None, None, None) == -1


def test_product_path_refuses_cpu_tensors():
    import torch
    from sinnerf_b200.nerf import NeRF, Embedding
    from sinnerf_b200.rendering import render_rays, sample_pdf
    emb = [Embedding(3, 10), Embedding(3, 4)]
    with pytest.raises(RuntimeError, match="no CPU path"):
        render_rays([NeRF(use_new_activation=True)], emb, torch.zeros(4, 8))
    with pytest.raises(RuntimeError, match="no CPU path"):
        emb[0](torch.zeros(4, 3))
    with pytest.raises(RuntimeError, match="no CPU path"):
        sample_pdf(torch.zeros(2, 5), torch.zeros(2, 4), 4, det=True)
    with pytest.raises(RuntimeError, match="no CPU path"):
        NeRF(use_new_activation=True)(torch.zeros(4, 90))


def test_state_dict_contract():
    """Parameter names/shapes are the reference's; seeded init equals the oracle's."""
    import torch
    from oracle.render_oracle import default_init_params, param_shapes
    from sinnerf_b200.nerf import NeRF
    torch.manual_seed(0)
    m = NeRF(use_new_activation=True)
    sd = m.state_dict()
    shapes = param_shapes()
    assert list(sd.keys()) == list(shapes.keys())
    for k, v in sd.items():
        assert tuple(v.shape) == shapes[k], k
    p = default_init_params(0)
    for k, v in sd.items():
        assert torch.equal(v, p[k]), k
    assert sum(v.numel() for v in sd.values()) == 595844
    # old-activation variant keeps the same keys
    assert list(NeRF().state_dict().keys()) == list(shapes.keys())
