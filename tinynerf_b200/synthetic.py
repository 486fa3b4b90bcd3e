"""Synthetic inputs shaped like the reference's workloads: packed-ray
microbench data, Blender-shaped camera rays, analytic occupancy grids.  Pure torch, seeded."""
from __future__ import annotations

import math
from typing import Tuple

import torch

DECAY = 0.01 ** (1 / 16)  # occupancy_grid_decay of the reference (src/run.py:109)


def packed_rays(n_samples: int, seed: int, mean_len: float = 64.0, max_len: int = 1024,
                empty_frac: float = 0.1, device="cpu") -> Tuple[torch.Tensor, torch.Tensor, torch.Tensor]:
    """Config 5: ray lengths 10% zero, else clip(Geometric(mean), 1, max_len), drawn until they sum to
    n_samples (last ray truncated).  -> (sigmas [N], info [R,2] int32, grad_weights [N]) on `device`;
    sigmas = exp(1.5*N(0,1) - 1)."""
    g = torch.Generator().manual_seed(seed)
    est = int(n_samples / (mean_len * (1 - empty_frac)) * 1.3) + 64
    lens = torch.zeros(0, dtype=torch.int64)
    while lens.sum() < n_samples:
        u = torch.rand(est, generator=g, dtype=torch.float64).clamp_min(1e-12)
        geo = torch.floor(torch.log(u) / math.log(1.0 - 1.0 / mean_len)).long() + 1
        geo = geo.clamp(1, max_len)
        geo[torch.rand(est, generator=g) < empty_frac] = 0
        lens = torch.cat([lens, geo])
    csum = torch.cumsum(lens, 0)
    last = int((csum >= n_samples).nonzero()[0])
    lens = lens[: last + 1].clone()
    lens[last] -= csum[last] - n_samples
    start = torch.cumsum(lens, 0) - lens
    info = torch.stack([start, lens], -1).to(torch.int32)
    sigmas = torch.exp(1.5 * torch.randn(n_samples, generator=g) - 1.0)
    grad = torch.randn(n_samples, generator=g)
    return sigmas.to(device), info.to(device), grad.to(device)


def blender_rays(n_rays: int, seed: int, radius: float = 4.0311, res: int = 800,
                 camera_angle_x: float = 0.6911112, device="cpu") -> Tuple[torch.Tensor, torch.Tensor]:
    """Random pixels of random cameras on the upper hemisphere looking at the origin, ray convention of
    src/data.py:53-69 (pixel centre +0.5, -fy, z=-1, d = grid @ R.T normalised)."""
    g = torch.Generator().manual_seed(seed)
    focal = 0.5 * res / math.tan(0.5 * camera_angle_x)
    theta = torch.rand(n_rays, generator=g) * 2 * math.pi
    phi = torch.rand(n_rays, generator=g) * 0.45 * math.pi + 0.05
    pos = radius * torch.stack([torch.cos(theta) * torch.cos(phi), torch.sin(theta) * torch.cos(phi), torch.sin(phi)], -1)
    fwd = -pos / pos.norm(dim=-1, keepdim=True)          # camera looks along -z
    up = torch.tensor([0.0, 0.0, 1.0]).expand_as(fwd)
    right = torch.linalg.cross(fwd, up)
    right = right / right.norm(dim=-1, keepdim=True)
    true_up = torch.linalg.cross(right, fwd)
    R = torch.stack([right, true_up, -fwd], -1)          # columns: x, y, z axes of the camera
    px = torch.rand(n_rays, 2, generator=g) * res
    px = torch.floor(px)
    gx = (px[:, 0] - res / 2 + 0.5) / focal
    gy = (px[:, 1] - res / 2 + 0.5) / (-focal)
    cam = torch.stack([gx, gy, -torch.ones_like(gx)], -1)
    d = torch.einsum("nij,nj->ni", R, cam)
    d = d / d.norm(dim=-1, keepdim=True)
    return pos.float().contiguous().to(device), d.float().contiguous().to(device)


def analytic_grid(res: int, seed: int, device="cpu") -> torch.Tensor:
    """Occupancy grid state of a trained scene: cells inside a ball (r=0.5) or a torus are 1, 4% of the
    other cells are "floaters" at decay^k, k ~ U{1..15} (above the 0.01 threshold), the rest have decayed
    below it (k ~ U{16..40}); about 10% of the cells are occupied."""
    g = torch.Generator().manual_seed(seed)
    ax = (torch.arange(res, dtype=torch.float32) + 0.5) / res * 3.0 - 1.5  # world coords of the aabb [-1.5,1.5]
    z, y, x = torch.meshgrid(ax, ax, ax, indexing="ij")
    ball = (x ** 2 + y ** 2 + z ** 2) < 0.5 ** 2
    torus = ((torch.sqrt(x ** 2 + y ** 2) - 0.9) ** 2 + z ** 2) < 0.25 ** 2
    floater = torch.rand(res, res, res, generator=g) < 0.04
    k_lo = torch.randint(1, 16, (res, res, res), generator=g).float()
    k_hi = torch.randint(16, 41, (res, res, res), generator=g).float()
    grid = torch.tensor(DECAY, dtype=torch.float32) ** torch.where(floater, k_lo, k_hi)
    grid[ball | torus] = 1.0
    return grid.to(device)


def analytic_density(x: torch.Tensor) -> torch.Tensor:
    """Density of the same ball+torus scene at contracted coords x in [-1,1] (aabb +-1.5)."""
    w = x * 1.5
    ball = (w ** 2).sum(-1) < 0.25
    torus = ((torch.sqrt(w[..., 0] ** 2 + w[..., 1] ** 2) - 0.9) ** 2 + w[..., 2] ** 2) < 0.0625
    return torch.where(ball | torus, torch.full_like(w[..., 0], 40.0), torch.zeros_like(w[..., 0]))


def _look_at(pos: torch.Tensor) -> torch.Tensor:
    """Camera-to-world rotations [n,3,3] (columns: right, up, -forward) of cameras at `pos` looking at the origin, z up."""
    fwd = -pos / pos.norm(dim=-1, keepdim=True)
    up = torch.tensor([0.0, 0.0, 1.0]).expand_as(fwd)
    right = torch.linalg.cross(fwd, up)
    right = right / right.norm(dim=-1, keepdim=True)
    true_up = torch.linalg.cross(right, fwd)
    return torch.stack([right, true_up, -fwd], -1)


def camera_rays(width: int, height: int, focal: float, pos, device="cpu") -> Tuple[torch.Tensor, torch.Tensor]:
    """Every pixel of ONE pinhole camera at `pos` looking at the origin, in the reference's ray convention and pixel order
    (src/data.py:49-73: meshgrid "xy", pixel centre +0.5, (fx, -fy), z = -1, d = grid @ R.T normalised).
    -> (rays_o, rays_d) [height*width, 3], row-major over (row, column) like PoseDataset's [h, w, 3] images."""
    pos = torch.as_tensor(pos, dtype=torch.float32)
    rot = _look_at(pos[None])[0]
    gx, gy = torch.meshgrid(torch.arange(width, dtype=torch.float32), torch.arange(height, dtype=torch.float32), indexing="xy")
    cam = torch.stack([(gx - width / 2 + 0.5) / focal, (gy - height / 2 + 0.5) / (-focal), -torch.ones_like(gx)], -1).view(-1, 3)
    d = cam @ rot.T
    d = d / d.norm(dim=-1, keepdim=True)
    return pos.expand_as(d).contiguous().to(device), d.contiguous().to(device)


def colmap_rays(n_rays: int, seed: int, n_cameras: int = 200, width: int = 1000, height: int = 750,
                focal: float = 800.0, device="cpu") -> Tuple[torch.Tensor, torch.Tensor, float]:
    """Config 4: random pixels of `n_cameras` inward-facing pinhole cameras on a jittered ring
    (radius 1.5 +- 0.2, height +- 0.3).  -> (rays_o, rays_d, scene_scale) with scene_scale = the largest per-axis variance
    of the camera positions, as NerfData.scene_scale computes it (src/data.py:75-76)."""
    g = torch.Generator().manual_seed(seed)
    ang = torch.rand(n_cameras, generator=g) * 2 * math.pi
    rad = 1.5 + (torch.rand(n_cameras, generator=g) * 2 - 1) * 0.2
    hgt = (torch.rand(n_cameras, generator=g) * 2 - 1) * 0.3
    cams = torch.stack([rad * torch.cos(ang), rad * torch.sin(ang), hgt], -1)
    rots = _look_at(cams)
    scene_scale = float(torch.max(torch.var(cams, 0)))
    which = torch.randint(0, n_cameras, (n_rays,), generator=g)
    px = torch.floor(torch.rand(n_rays, generator=g) * width)
    py = torch.floor(torch.rand(n_rays, generator=g) * height)
    cam = torch.stack([(px - width / 2 + 0.5) / focal, (py - height / 2 + 0.5) / (-focal), -torch.ones_like(px)], -1)
    d = torch.einsum("nij,nj->ni", rots[which], cam)
    d = d / d.norm(dim=-1, keepdim=True)
    return cams[which].float().contiguous().to(device), d.float().contiguous().to(device), scene_scale
