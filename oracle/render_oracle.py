"""CPU oracle for the SinNeRF volumetric-rendering hot path.

TEST INFRASTRUCTURE ONLY.  This file is a stage-by-stage restatement, in plain
torch-on-CPU arithmetic, of the algorithm in the reference's
``models/rendering.py`` / ``models/nerf.py`` / ``models/activations.py``.
Only ``tests/``, ``__graft_entry__.smoke()`` and the ``cpu_baseline`` /
``--impl reference`` legs of ``bench.py`` may import it.  The product package
``sinnerf_b200`` never does: its entry points raise when the CUDA library is
missing instead of falling back to anything here.

Parity pin: the reference has no tests or golden vectors of its own for this
path, so the oracle is pinned against outputs of the
reference itself, generated in the build container by
``tests/golden/make_golden.py`` (which imports ``/root/reference``) and
committed as ``tests/golden/*.npz``.  ``tests/test_oracle_golden.py`` checks
every stage of this file against those fixtures and against the known-answer
vectors of SURVEY.md section 8c.

Every function cites the reference lines it follows (paths relative to the
reference checkout).  Weights are passed as a flat ``{name: tensor}`` dict with
the reference's state-dict key names (``xyz_encoding_1.0.weight`` ...), so the
oracle does not depend on any module class.
"""
from __future__ import annotations

import math
from typing import Dict, Optional

import torch

Tensor = torch.Tensor
Params = Dict[str, Tensor]

N_XYZ_FREQS = 10   # Embedding(3, 10)  -> 63 channels  (eval.py:134, sinnerf.py:131)
N_DIR_FREQS = 4    # Embedding(3, 4)   -> 27 channels  (eval.py:135, sinnerf.py:132)


# --------------------------------------------------------------------------- #
# a-4  positional encoding                                                     #
# --------------------------------------------------------------------------- #
def embed(x: Tensor, n_freqs: int) -> Tensor:
    """models/nerf.py:24-41 with freq_bands = 2**linspace(0, L-1, L) (nerf.py:19-20).

    Output channel order: [x, sin(1x), cos(1x), sin(2x), cos(2x), ...], each
    block ``x.shape[-1]`` wide.  ``2**k * x`` is exact in binary floating point.
    """
    blocks = [x]
    for k in range(n_freqs):
        xk = x * float(2 ** k)
        blocks.append(torch.sin(xk))
        blocks.append(torch.cos(xk))
    return torch.cat(blocks, dim=-1)


# --------------------------------------------------------------------------- #
# a-12  activations                                                            #
# --------------------------------------------------------------------------- #
def shifted_softplus(x: Tensor) -> Tensor:
    """models/activations.py:23-35: softplus(x-1) in the overflow-safe form."""
    s = x - 1
    return torch.log1p(torch.exp(-s.abs())) + s * (s >= 0)


def widened_sigmoid(x: Tensor) -> Tensor:
    """models/activations.py:8-20: 0.5*(1 + (1+2e-3)*tanh(x/2))."""
    return 0.5 * (1.0 + (1.0 + 2.0 * 1e-3) * torch.tanh(0.5 * x))


# --------------------------------------------------------------------------- #
# a-6  the 8x256 field MLP                                                     #
# --------------------------------------------------------------------------- #
def _round_st(x: Tensor, dt) -> Tensor:
    """x rounded to dtype `dt` in the forward, identity in the backward (None: x itself)."""
    return x if dt is None else x + (x.to(dt).to(x.dtype) - x).detach()


def _affine(p: Params, name: str, x: Tensor, linear_dtype=None) -> Tensor:
    """nn.Linear.  linear_dtype (e.g. torch.bfloat16): BOTH operands rounded to it, fp32 accumulate -- the arithmetic
    of the product's reduced-precision modes (the reference itself has no such path; SURVEY 8c defines the bf16
    configuration as 'nn.Linear in half precision, everything else fp32', i.e. autocast)."""
    return torch.addmm(p[name + ".bias"], _round_st(x, linear_dtype), _round_st(p[name + ".weight"], linear_dtype).t())


def field_mlp(p: Params, xyz_enc: Tensor, dir_enc: Optional[Tensor],
              sigma_only: bool = False, new_activation: bool = True,
              depth: int = 8, skips=(4,), linear_dtype=None, fold_bottleneck: bool = False) -> Tensor:
    """models/nerf.py:105-148.

    xyz_enc (P,63), dir_enc (P,27) -> (P,4) = [r,g,b,sigma]  (or (P,1) sigma).
    Skip concat order is [input_xyz, hidden] (nerf.py:132-133); the direction
    layer sees [bottleneck, dir] (nerf.py:142); sigma and the bottleneck have no
    activation (nerf.py:136,140).
    """
    h = xyz_enc
    for i in range(depth):
        if i in skips:
            h = torch.cat([xyz_enc, h], dim=-1)
        h = torch.relu_(_affine(p, f"xyz_encoding_{i + 1}.0", h, linear_dtype))  # nn.ReLU(True), nerf.py:73
    sigma = _affine(p, "sigma", h)          # the heads are fp32 dot products in every mode
    if sigma_only:
        return sigma
    if fold_bottleneck:
        # the product's tensor-core modes fold the activation-free bottleneck into the direction layer:
        # Wd[:, :W] (Wf h + bf) = (Wd[:, :W] Wf) h + Wd[:, :W] bf, the product formed in fp64 and rounded once
        Wd, bd = p["dir_encoding.0.weight"], p["dir_encoding.0.bias"]
        Wf, bf = p["xyz_encoding_final.weight"], p["xyz_encoding_final.bias"]
        w = h.shape[-1]
        Wp = (Wd[:, :w].double() @ Wf.double()).to(h.dtype)
        bp = bd + (Wd[:, :w].double() @ bf.double()).to(h.dtype)
        g = bp + _round_st(h, linear_dtype) @ _round_st(Wp, linear_dtype).t() \
            + _round_st(dir_enc, linear_dtype) @ _round_st(Wd[:, w:], linear_dtype).t()
    else:
        feat = _affine(p, "xyz_encoding_final", h, linear_dtype)
        g = _affine(p, "dir_encoding.0", torch.cat([feat, dir_enc], dim=-1), linear_dtype)
    g = shifted_softplus(g) if new_activation else torch.relu(g)
    c = _affine(p, "rgb.0", g)
    c = widened_sigmoid(c) if new_activation else torch.sigmoid(c)
    return torch.cat([c, sigma], dim=-1)


# --------------------------------------------------------------------------- #
# a-2  stratified depth sampling                                               #
# --------------------------------------------------------------------------- #
def sample_z(near: Tensor, far: Tensor, n_samples: int, use_disp: bool = False,
             perturb: float = 0.0, perturb_u: Optional[Tensor] = None) -> Tensor:
    """models/rendering.py:264-282.

    near, far (N,1).  ``z_steps`` is torch.linspace in the *default* dtype
    (rendering.py:264 passes no dtype).  perturb_u is the U[0,1) tensor the
    reference draws at :281; pass it to replay a specific draw.
    """
    t = torch.linspace(0, 1, n_samples).to(near)       # values formed on the host (as the reference does), then moved
    if not use_disp:
        z = near * (1 - t) + far * t
    else:
        z = 1 / (1 / near * (1 - t) + 1 / far * t)
    z = z.expand(near.shape[0], n_samples)
    if perturb > 0:
        mid = 0.5 * (z[:, :-1] + z[:, 1:])
        upper = torch.cat([mid, z[:, -1:]], dim=-1)
        lower = torch.cat([z[:, :1], mid], dim=-1)
        if perturb_u is None:
            perturb_u = torch.rand(z.shape, device=z.device)
        z = lower + (upper - lower) * (perturb * perturb_u)
    return z


# --------------------------------------------------------------------------- #
# a-7  sigma -> alpha -> transmittance compositing                             #
# --------------------------------------------------------------------------- #
def composite(sigma: Tensor, z: Tensor, d_norm: Tensor, rgb: Optional[Tensor] = None,
              noise: Optional[Tensor] = None, white_back: bool = False):
    """models/rendering.py:215-248.

    sigma (N,S), z (N,S), d_norm (N,1) = ||rays_d||, rgb (N,S,3) or None,
    noise (N,S) = randn*noise_std already scaled, or None for zero.
    Returns weights (N,S) alone if rgb is None (the weights_only branch,
    :237-238), else (rgb_map (N,3), depth_map (N,), weights (N,S)).
    """
    delta = z[:, 1:] - z[:, :-1]
    delta = torch.cat([delta, torch.full_like(delta[:, :1], 1e10)], dim=-1)
    delta = delta * d_norm
    s = sigma if noise is None else sigma + noise
    alpha = 1 - torch.exp(-delta * torch.relu(s))
    shifted = torch.cat([torch.ones_like(alpha[:, :1]), 1 - alpha + 1e-10], dim=-1)
    weights = alpha * torch.cumprod(shifted, dim=-1)[:, :-1]
    if rgb is None:
        return weights
    rgb_map = (weights.unsqueeze(-1) * rgb).sum(dim=-2)
    depth_map = (weights * z).sum(dim=-1)
    if white_back:
        rgb_map = rgb_map + 1 - weights.sum(dim=1).unsqueeze(-1)
    return rgb_map, depth_map, weights


# --------------------------------------------------------------------------- #
# a-8  inverse-CDF importance sampling                                         #
# --------------------------------------------------------------------------- #
def sample_pdf(bins: Tensor, weights: Tensor, n_importance: int, det: bool = False,
               eps: float = 1e-5, u: Optional[Tensor] = None) -> Tensor:
    """models/rendering.py:15-61.

    bins (N,M+1), weights (N,M) -> (N,n_importance).  ``searchsorted(right=True)``
    returns #{j : cdf_j <= u}.  ``u`` replays the rand draw at :43 when det is False.
    """
    n_rays, m = weights.shape
    w = weights + eps
    pdf = w / w.sum(dim=-1, keepdim=True)
    cdf = torch.cat([torch.zeros_like(pdf[:, :1]), torch.cumsum(pdf, dim=-1)], dim=-1)
    if det:
        u = torch.linspace(0, 1, n_importance).to(bins).expand(n_rays, n_importance)
    elif u is None:
        u = torch.rand(n_rays, n_importance, device=bins.device).to(bins.dtype)
    u = u.contiguous()
    idx = torch.searchsorted(cdf, u, right=True)
    lo = (idx - 1).clamp_min(0)
    hi = idx.clamp_max(m)
    cdf_lo, cdf_hi = torch.gather(cdf, 1, lo), torch.gather(cdf, 1, hi)
    bin_lo, bin_hi = torch.gather(bins, 1, lo), torch.gather(bins, 1, hi)
    denom = cdf_hi - cdf_lo
    denom = torch.where(denom < eps, torch.ones_like(denom), denom)
    return bin_lo + (u - cdf_lo) / denom * (bin_hi - bin_lo)


# --------------------------------------------------------------------------- #
# a-3 + a-5 + a-6 + a-7  one pass of `inference`                               #
# --------------------------------------------------------------------------- #
def field_pass(p: Params, rays_o: Tensor, rays_d: Tensor, dir_enc: Tensor, z: Tensor,
               noise: Optional[Tensor], white_back: bool, weights_only: bool = False,
               new_activation: bool = True, point_chunk: int = 1 << 15, linear_dtype=None,
               fold_bottleneck: bool = False):
    """models/rendering.py:161-248 (the nested ``inference``) incl. the point
    generation at :284-285 / :317-318.  Returns a dict with raw (N,S,4) (or
    sigma (N,S)), rgb, depth, weights.
    """
    n, s = z.shape
    xyz = (rays_o.unsqueeze(1) + rays_d.unsqueeze(1) * z.unsqueeze(2)).reshape(-1, 3)
    dirs = None if weights_only else torch.repeat_interleave(dir_enc, repeats=s, dim=0)
    outs = []
    for i in range(0, xyz.shape[0], point_chunk):
        e = embed(xyz[i:i + point_chunk], N_XYZ_FREQS)
        outs.append(field_mlp(p, e, None if weights_only else dirs[i:i + point_chunk],
                              sigma_only=weights_only, new_activation=new_activation, linear_dtype=linear_dtype,
                              fold_bottleneck=fold_bottleneck))
    raw = torch.cat(outs, dim=0)
    d_norm = torch.norm(rays_d.unsqueeze(1), dim=-1)
    if weights_only:
        sigma = raw.view(n, s)
        return {"sigma": sigma, "weights": composite(sigma, z, d_norm, None, noise, white_back)}
    raw = raw.view(n, s, 4)
    rgb_map, depth_map, weights = composite(raw[..., 3], z, d_norm, raw[..., :3], noise, white_back)
    return {"raw": raw, "rgb": rgb_map, "depth": depth_map, "weights": weights}


# --------------------------------------------------------------------------- #
# whole render_rays                                                            #
# --------------------------------------------------------------------------- #
def render_rays(coarse: Params, fine: Optional[Params], rays: Tensor, N_samples: int = 64,
                use_disp: bool = False, perturb: float = 0.0, noise_std: float = 1.0,
                N_importance: int = 0, white_back: bool = False, test_time: bool = False,
                new_activation: bool = True, rng: Optional[Dict[str, Tensor]] = None,
                z_fine_override: Optional[Tensor] = None, return_intermediates: bool = False,
                linear_dtype=None, fold_bottleneck: bool = False):
    """models/rendering.py:126-335.

    linear_dtype / fold_bottleneck (not in the reference): restate the product's reduced-precision MLP modes --
    see _affine / field_mlp -- so that those modes have a tight oracle of their own.

    ``rng`` may hold the four random tensors the reference draws, in its order
    (SURVEY.md 8a): 'perturb_u' (N,Sc) rand, 'noise_coarse' (N,Sc) randn,
    'pdf_u' (N,Ni) rand, 'noise_fine' (N,Sf) randn.  Missing entries are drawn
    here from torch's global CPU generator in that same order, so with
    ``rng=None`` and the same seed the result equals the reference on CPU
    bit for bit.  z_fine_override injects fine-pass sample depths (stage-wise
    parity of the chaotic fine pass, SURVEY.md hard part 3).
    """
    rng = dict(rng or {})
    n = rays.shape[0]
    rays_o, rays_d = rays[:, 0:3], rays[:, 3:6]
    near, far = rays[:, 6:7], rays[:, 7:8]
    dir_enc = embed(rays_d, N_DIR_FREQS)                       # rendering.py:261

    if perturb > 0 and "perturb_u" not in rng:
        rng["perturb_u"] = torch.rand(n, N_samples, device=rays.device)            # RNG call 1 (:281)
    z = sample_z(near, far, N_samples, use_disp, perturb, rng.get("perturb_u"))

    if "noise_coarse" not in rng:
        rng["noise_coarse"] = torch.randn(n, N_samples, device=rays.device)        # RNG call 2 (:224)
    noise_c = rng["noise_coarse"].to(rays.dtype) * noise_std
    c = field_pass(coarse, rays_o, rays_d, dir_enc, z, noise_c, white_back,
                   weights_only=test_time, new_activation=new_activation, linear_dtype=linear_dtype,
                   fold_bottleneck=fold_bottleneck)
    out = {"opacity_coarse": c["weights"]}
    if not test_time:
        out["rgb_coarse"], out["depth_coarse"] = c["rgb"], c["depth"]
    inter = {"z_coarse": z, "raw_coarse": c.get("raw", c.get("sigma"))}

    if N_importance > 0:
        z_mid = 0.5 * (z[:, :-1] + z[:, 1:])                    # :310
        det = perturb == 0
        if not det and "pdf_u" not in rng:
            rng["pdf_u"] = torch.rand(n, N_importance, device=rays.device)         # RNG call 3 (:43)
        z_new = sample_pdf(z_mid, c["weights"][:, 1:-1], N_importance, det=det,
                           u=None if det else rng["pdf_u"].to(rays.dtype)).detach()
        # ^ detach: no gradient flows from the fine pass into the coarse weights (:311-313)
        z_f, _ = torch.sort(torch.cat([z, z_new], dim=-1), dim=-1)   # :315
        if z_fine_override is not None:
            z_f = z_fine_override
        if "noise_fine" not in rng:
            rng["noise_fine"] = torch.randn(n, N_samples + N_importance, device=rays.device)  # RNG call 4
        noise_f = rng["noise_fine"].to(rays.dtype) * noise_std
        f = field_pass(fine, rays_o, rays_d, dir_enc, z_f, noise_f, white_back,
                       new_activation=new_activation, linear_dtype=linear_dtype, fold_bottleneck=fold_bottleneck)
        out["rgb_fine"], out["depth_fine"], out["opacity_fine"] = f["rgb"], f["depth"], f["weights"]
        inter.update(z_new=z_new, z_fine=z_f, raw_fine=f["raw"])
    else:
        # rendering.py:330-333 (raises UnboundLocalError under test_time in the
        # reference; mirrored as KeyError here)
        out["rgb_fine"], out["depth_fine"] = out["rgb_coarse"], out["depth_coarse"]
        out["opacity_fine"] = out["opacity_coarse"]
    if return_intermediates:
        out["_inter"] = inter
    return out


# --------------------------------------------------------------------------- #
# helpers shared by tests / bench (not part of the reference algorithm)        #
# --------------------------------------------------------------------------- #
def param_shapes(depth: int = 8, width: int = 256, c_xyz: int = 63, c_dir: int = 27, skips=(4,)):
    """State-dict key -> shape, in nn.Module registration order (nerf.py:66-103)."""
    shapes = {}
    for i in range(depth):
        k = c_xyz if i == 0 else (width + c_xyz if i in skips else width)
        shapes[f"xyz_encoding_{i + 1}.0.weight"] = (width, k)
        shapes[f"xyz_encoding_{i + 1}.0.bias"] = (width,)
    shapes["xyz_encoding_final.weight"] = (width, width)
    shapes["xyz_encoding_final.bias"] = (width,)
    shapes["dir_encoding.0.weight"] = (width // 2, width + c_dir)
    shapes["dir_encoding.0.bias"] = (width // 2,)
    shapes["sigma.weight"] = (1, width)
    shapes["sigma.bias"] = (1,)
    shapes["rgb.0.weight"] = (3, width // 2)
    shapes["rgb.0.bias"] = (3,)
    return shapes


def default_init_params(seed: int, dtype=torch.float32) -> Params:
    """Same numbers as ``torch.manual_seed(seed); NeRF(use_new_activation=True)``
    in the reference: nn.Linear default init (kaiming_uniform(a=sqrt(5)) on the
    weight, then U(-1/sqrt(fan_in), 1/sqrt(fan_in)) on the bias), layers created
    in the order of nerf.py:66-103.  Checked against the reference module in
    tests/golden/make_golden.py.
    """
    gen_state = torch.random.get_rng_state()
    torch.manual_seed(seed)
    p: Params = {}
    shapes = param_shapes()
    names = [k[:-len(".weight")] for k in shapes if k.endswith(".weight")]
    for name in names:
        out_f, in_f = shapes[name + ".weight"]
        w = torch.empty(out_f, in_f)
        torch.nn.init.kaiming_uniform_(w, a=math.sqrt(5))
        bound = 1 / math.sqrt(in_f)
        b = torch.empty(out_f).uniform_(-bound, bound)
        p[name + ".weight"], p[name + ".bias"] = w.to(dtype), b.to(dtype)
    torch.random.set_rng_state(gen_state)
    return p


# --------------------------------------------------------------------------- #
# f-1  ray generation (the step before the path)                               #
# --------------------------------------------------------------------------- #
def camera_rays(H: int, W: int, focal, c2w: Tensor, near: float, far: float, center=None, opencv: bool = False,
                window=None) -> Tensor:
    """datasets/ray_utils.py:73-120 (get_ray_directions + get_rays) / datasets/dtu_proj.py:17-34, and the
    [o, d, near, far] concatenation the datasets do.  window = (row0, col0, rows, cols, stride)."""
    fx, fy = (focal, focal) if not isinstance(focal, (tuple, list)) else focal
    cx, cy = (W / 2, H / 2) if center is None else center
    row0, col0, rows, cols, stride = (0, 0, H, W, 1) if window is None else window
    jj = torch.arange(rows, dtype=torch.float32) * stride + row0
    ii = torch.arange(cols, dtype=torch.float32) * stride + col0
    j, i = torch.meshgrid(jj, ii, indexing="ij")
    if opencv:
        d = torch.stack([(i - cx) / fx, (j - cy) / fy, torch.ones_like(i)], -1)
    else:
        d = torch.stack([(i - cx) / fx, -(j - cy) / fy, -torch.ones_like(i)], -1)
    d = (d @ c2w[:, :3].T).reshape(-1, 3)
    o = c2w[:, 3].expand(d.shape)
    n = d.shape[0]
    return torch.cat([o, d, torch.full((n, 1), float(near)), torch.full((n, 1), float(far))], dim=-1)
