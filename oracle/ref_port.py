"""Functional PyTorch restatement of the reference's hot path.  TEST INFRASTRUCTURE ONLY.

Every function states, in plain un-fused torch ops, what the cited reference lines compute
(loicmagne/tinynerf, paths relative to its root).  It is device-agnostic: on CPU it is the
"pure-PyTorch CPU path" used as cpu_baseline / --impl reference; on the GPU box the very same code run
with device="cuda" is the reference-equivalent oracle for bit-exact mask/packing parity (the
reference's arithmetic lives in torch's own kernels, which exist on the box; /root/reference does not).

Pinned against the real reference by tests/golden/make_golden.py (imports /root/reference/src) and
tests/test_oracle_golden.py.  The weights op has no PyTorch implementation in the reference
(src/cuda.cu is CUDA-only); `weights_op` dispatches to oracle/tnf_oracle.c on CPU and to the
unmodified reference kernel in oracle/_ref/_cuda.so on CUDA.
"""
from __future__ import annotations

import math
from typing import Callable, List, Sequence, Tuple

import torch
import torch.nn.functional as F

Tensor = torch.Tensor

# ---- contraction (src/core.py:11-33) -----------------------------------------------------------


def contract_mip360(p: Tensor, order: float = math.inf) -> Tensor:
    """src/core.py:18-19"""
    n = torch.norm(p, p=order, dim=-1, keepdim=True)
    return torch.where(n <= 1.0, p, (2.0 - 1.0 / n) * p / n) / 2.0


def contract_aabb(p: Tensor, aabb: Tensor) -> Tuple[Tensor, Tensor]:
    """src/core.py:29-30 -> (coords in [-1,1], inside mask)"""
    lo, hi = aabb[0], aabb[1]
    inside = ((p >= lo) & (p <= hi)).all(dim=-1)
    return (p - lo) / (hi - lo) * 2.0 - 1.0, inside


# ---- ray marching (src/core.py:36-90) ----------------------------------------------------------


def unbounded_tables(n_samples: int, near: float, uniform_range: float, device) -> Tuple[Tensor, Tensor]:
    """src/core.py:52-55: ray independent t values [S] and step sizes [S]"""
    x = torch.linspace(0.0, 1.0 - (1.0 / (n_samples + 2)), n_samples + 1, device=device)
    t = torch.where(x < 0.5, 2 * x, 1 / (2 - 2 * x)) * uniform_range + near
    return t[:-1], t[1:] - t[:-1]


def aabb_step_size(aabb: Tensor, n_samples: int) -> Tensor:
    """src/core.py:68-70 (0-dim tensor on aabb's device)"""
    return torch.norm(aabb[1] - aabb[0]) / n_samples


def aabb_t_min(rays_o: Tensor, rays_d: Tensor, aabb: Tensor, near: float, far: float) -> Tensor:
    """src/core.py:77-81"""
    dist = aabb.unsqueeze(1) - rays_o
    hit = dist / torch.where(rays_d == 0.0, rays_d + 1e-9, rays_d)
    return torch.clamp(torch.amax(torch.amin(hit, dim=0), dim=1), min=near, max=far)


# ---- occupancy grid (src/core.py:93-156) -------------------------------------------------------


def occupancy_values(grid: Tensor, coords: Tensor) -> Tensor:
    """src/core.py:151-155: trilinear lookup, coords[...,0] -> last grid dim"""
    shape = coords.shape[:-1]
    return F.grid_sample(grid[None, None], coords.reshape(1, -1, 1, 1, 3), align_corners=True).view(shape)


def occupancy_lookup(grid: Tensor, coords: Tensor, threshold: float) -> Tensor:
    """src/core.py:156"""
    return occupancy_values(grid, coords) > threshold


def occupancy_threshold(base: float, mean: float) -> float:
    """src/core.py:126-127"""
    return min(base, mean)


def occupancy_update(grid: Tensor, noise: Tensor, sigma_fn: Callable[[Tensor], Tensor], step_size,
                     threshold: float, decay: float) -> Tuple[Tensor, float]:
    """src/core.py:134-145 with the per-slice jitter supplied as noise[D,H,W,3] (the reference draws
    it from the CPU generator, :137).  Returns (new grid, new mean)."""
    D, H, W = grid.shape
    size = torch.tensor([D, H, W], dtype=torch.float)
    zz, yy, xx = torch.meshgrid(torch.arange(D, dtype=torch.float), torch.arange(H, dtype=torch.float),
                                torch.arange(W, dtype=torch.float), indexing="ij")
    cells = torch.stack([xx, yy, zz], -1)  # flipped (x,y,z), src/core.py:119
    out = grid.clone()
    for i in range(D):
        c = -1.0 + 2.0 * (cells[i] + noise[i].cpu()) / size
        c = c.view(-1, 3).to(grid.device)
        alpha = (1.0 - torch.exp(-sigma_fn(c) * step_size)).view(H, W)
        out[i] = torch.where(alpha > threshold, torch.ones_like(alpha), decay * out[i])
    return out, out.mean().item()


# ---- ray provider (src/core.py:158-188) --------------------------------------------------------


def ray_provider(rays_o: Tensor, rays_d: Tensor, grid: Tensor, threshold: float, *, scene: str, n_samples: int,
                 aabb: Tensor | None = None, near: float = 0.0, far: float = 1e5, uniform_range: float = 1.0,
                 noise: Tensor | None = None):
    """-> (packed [N,7], info [R,2] int32, mask [R,S]).  noise=None means training=False."""
    R = rays_o.size(0)
    if scene == "aabb":
        step = aabb_step_size(aabb, n_samples)
        t_min = aabb_t_min(rays_o, rays_d, aabb, near, far)
        t = t_min[:, None] + torch.arange(n_samples, dtype=torch.float, device=rays_o.device) * step
        steps = torch.full_like(t, step)
    else:
        tt, ss = unbounded_tables(n_samples, near, uniform_range, rays_o.device)
        t, steps = tt.expand(R, n_samples), ss.expand(R, n_samples)
    if noise is not None:
        t = t + noise * steps                                      # :173
    pts = rays_o[:, None, :] + rays_d[:, None, :] * t[..., None]   # :174
    if scene == "aabb":
        pts, inside = contract_aabb(pts, aabb)
        mask = inside & occupancy_lookup(grid, pts, threshold)     # :176
    else:
        pts = contract_mip360(pts)
        mask = occupancy_lookup(grid, pts, threshold)
    count = mask.sum(dim=-1, dtype=torch.int)                      # :179-181
    start = torch.cumsum(count, dim=0, dtype=torch.int) - count
    info = torch.stack([start, count], -1)
    packed = torch.cat([pts[mask], torch.repeat_interleave(rays_d, count, 0), steps[mask][:, None]], -1)
    return packed, info, mask


# ---- weights op (src/core.py:192-207 over src/cuda.cu) -----------------------------------------


class _Weights(torch.autograd.Function):
    @staticmethod
    def forward(ctx, sigmas, steps, info, threshold):
        sigmas, steps, info = sigmas.contiguous(), steps.contiguous(), info.contiguous()
        if sigmas.is_cuda:
            from . import load_ref_cuda
            ref = load_ref_cuda()
            if ref is None:
                raise RuntimeError("oracle/_ref/_cuda.so missing (python oracle/build_ref.py)")
            if info.size(0) > (1 << 20):
                raise RuntimeError("reference kernel is only valid for <= 2^20 rays (swapped launch, src/cuda.cu:81-86)")
            w = ref.compute_weights_fwd(sigmas, steps, info, threshold)
        else:
            from . import c
            w = c.weights_fwd(sigmas, steps, info, threshold)
        ctx.save_for_backward(sigmas, steps, info, w)
        return w

    @staticmethod
    def backward(ctx, g):
        sigmas, steps, info, w = ctx.saved_tensors
        g = g.contiguous()
        if sigmas.is_cuda:
            from . import load_ref_cuda
            gs = load_ref_cuda().compute_weights_bwd(sigmas, steps, info, w, g)
        else:
            from . import c
            gs = c.weights_bwd(sigmas, steps, info, w, g)
        return gs, None, None, None


def weights_op(sigmas: Tensor, steps: Tensor, info: Tensor, threshold: float) -> Tensor:
    return _Weights.apply(sigmas, steps, info, threshold)


# ---- models (src/models.py) --------------------------------------------------------------------


def positional_encoding(x: Tensor, n_freqs: int) -> Tensor:
    """src/models.py:33-39: per coordinate [sin(2^k pi x)]_k then [cos(2^k pi x)]_k"""
    freqs = (2 ** torch.arange(0, n_freqs) * torch.pi).to(x.device)
    y = x[..., None] * freqs
    return torch.cat([torch.sin(y), torch.cos(y)], -1).flatten(-2)


def mlp(layers: Sequence[Tuple[Tensor, Tensor]], x: Tensor) -> Tensor:
    """src/models.py:17-28: Linear, ReLU, ..., Linear (no activation after the last)"""
    for i, (w, b) in enumerate(layers):
        x = F.linear(x, w, b)
        if i + 1 < len(layers):
            x = torch.relu(x)
    return x


class _TruncExp(torch.autograd.Function):
    """src/models.py:42-53"""

    @staticmethod
    def forward(ctx, x):
        ctx.save_for_backward(x)
        return torch.exp(x)

    @staticmethod
    def backward(ctx, g):
        return g * torch.exp(torch.clamp(ctx.saved_tensors[0], min=-15, max=15))


def sigma_head(layers, feats: Tensor) -> Tensor:
    """VanillaOpacityDecoder, src/models.py:70-77"""
    return _TruncExp.apply(mlp(layers, feats) - 1.0)


def rgb_head(layers, n_freqs: int, feats: Tensor, dirs: Tensor) -> Tensor:
    """VanillaColorDecoder, src/models.py:79-89"""
    return torch.sigmoid(mlp(layers, torch.cat([positional_encoding(dirs, n_freqs), dirs, feats], -1)))


def plane_lookup(plane: Tensor, xy: Tensor) -> Tensor:
    """KPlanesFeaturePlane.forward, src/models.py:105-113. plane [1,C,H,W], xy [N,2] -> [N,C]"""
    out = F.grid_sample(plane, xy.reshape(1, -1, 1, 2), align_corners=True)
    return out.squeeze().transpose(0, -1).contiguous().view(*xy.shape[:-1], plane.shape[1])


def kplanes_features(planes: Sequence[Sequence[Tensor]], x: Tensor) -> Tensor:
    """KPlanesFeatureField.forward, src/models.py:153-163. planes[scale][pair], pairs (0,1),(0,2),(1,2)"""
    pairs = [(0, 1), (0, 2), (1, 2)]
    feats = []
    for scale in planes:
        cur = 1.0
        for (i, j), plane in zip(pairs, scale):
            cur = cur * plane_lookup(plane, x[..., (i, j)])
        feats.append(cur)
    return torch.cat(feats, -1)


def kplanes_tv(planes: Sequence[Sequence[Tensor]]) -> Tensor:
    """KPlanesFeatureField.loss_tv, src/models.py:115-118,165-172"""
    tot, cnt = 0.0, 0
    for scale in planes:
        for p in scale:
            tot = tot + F.mse_loss(p[:, :, 1:, :], p[:, :, :-1, :]) + F.mse_loss(p[:, :, :, 1:], p[:, :, :, :-1])
            cnt += 1
    return tot / cnt


def grid3_lookup(grid: Tensor, x: Tensor) -> Tensor:
    """CobafaGrid.forward, src/models.py:228-238. grid [1,C,D,H,W], x [N,3] -> [N,C]"""
    out = F.grid_sample(grid, x.reshape(1, -1, 1, 1, 3), align_corners=True)
    return out.squeeze().transpose(0, -1).contiguous().view(*x.shape[:-1], grid.shape[1])


def cobafa_lookup(basis: Sequence[Tensor], coef: Tensor, freqs: Sequence[float], x: Tensor) -> Tensor:
    """CobafaFeatureField.forward up to the concat, src/models.py:259-264"""
    c = grid3_lookup(coef, x)
    ys = []
    for i, (f, b) in enumerate(zip(freqs, basis)):
        enc = 2.0 * ((f * x) % 1.0) - 1.0  # SawtoothEncoding, src/models.py:213-214
        ys.append(grid3_lookup(b, enc) * c[:, [i]])
    return torch.cat(ys, -1)


# ---- renderer (src/core.py:225-267) ------------------------------------------------------------


def render(feature_fn: Callable[[Tensor], Tensor], sigma_fn: Callable[[Tensor], Tensor],
           rgb_fn: Callable[[Tensor, Tensor], Tensor], packed: Tensor, info: Tensor, bg: Tensor | None,
           threshold: float = 1e-4, return_aux: bool = False):
    """NerfRenderer.forward for a non-empty batch."""
    n, r = packed.size(0), info.size(0)
    feats = feature_fn(packed[:, :3])
    sigmas = sigma_fn(feats).ravel()
    w = weights_op(sigmas, packed[:, 6], info, threshold)
    m = w > 0.0
    rgbs = torch.zeros((n, 3), device=packed.device)
    rgbs[m] = rgb_fn(feats[m], packed[:, 3:6][m])
    rgbs = rgbs * w[:, None]
    out = torch.zeros((r, 3), device=packed.device)
    ray_id = torch.repeat_interleave(torch.arange(r, device=packed.device), info[:, 1])
    out.index_add_(0, ray_id, rgbs)
    if bg is not None:
        op = torch.zeros(r, device=packed.device)
        op.index_add_(0, ray_id, w)
        out = out + bg.to(packed.device) * (1 - op[:, None])
    if return_aux:
        return out, dict(features=feats, sigmas=sigmas, weights=w)
    return out
