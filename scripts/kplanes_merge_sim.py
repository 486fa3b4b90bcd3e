"""Diagnostics (CPU, numpy/torch only): how many corner lines of the K-Planes scatter repeat among G consecutive packed
samples -- the most an in-ray merge of the reductions could save.  Rays of the bench scene are marched inline (AABB slab
entry, 256 uniform steps with jitter, trilinear occupancy lookup of the analytic grid), kept samples in (ray, step) order.
Output under profiles/r02_kplanes_merge_sim.txt."""
import sys
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
import numpy as np
import torch
import torch.nn.functional as F
from tinynerf_b200 import synthetic

torch.manual_seed(0)
R, S = 10240, 256
o, d = synthetic.blender_rays(R, seed=5)
grid = synthetic.analytic_grid(128, seed=3)
lo, hi = torch.full((3,), -1.5), torch.full((3,), 1.5)
step = float((hi - lo).norm()) / S
dd = torch.where(d == 0, torch.full_like(d, 1e-9), d)
t0, t1 = (lo - o) / dd, (hi - o) / dd
t_min = torch.minimum(t0, t1).max(-1).values.clamp(0.1, 1e5)
t = t_min[:, None] + (torch.arange(S) + torch.rand(R, S)) * step
pts = o[:, None, :] + d[:, None, :] * t[..., None]
inside = ((pts >= lo) & (pts <= hi)).all(-1)
x = (pts - lo) / (hi - lo) * 2 - 1
occ = F.grid_sample(grid[None, None], x.view(1, -1, 1, 1, 3), align_corners=True).view(R, S)
keep = inside & (occ > 0.01)
xs = x[keep].numpy().astype(np.float64)          # row-major (ray, step) order = the packed order
N = len(xs)
print("samples", N, "rays with samples", int(keep.any(-1).sum()))
for res in (128, 256, 512):
    tot, out, runs_cell = 0, {G: 0 for G in (4, 8, 16, 32)}, 0
    for a, b in ((0, 1), (0, 2), (1, 2)):
        iu, iv = (xs[:, a] + 1) * 0.5 * (res - 1), (xs[:, b] + 1) * 0.5 * (res - 1)
        cell = np.floor(iv).astype(np.int64) * res + np.floor(iu).astype(np.int64)
        corners = np.stack([cell, cell + 1, cell + res, cell + res + 1], 1)
        tot += 4 * N
        for G in out:
            ng = N // G
            c = np.sort(corners[:ng * G].reshape(ng, G * 4), axis=1)
            out[G] += int((1 + (np.diff(c, axis=1) != 0).sum(1)).sum())
        runs_cell += N - int((cell[1:] == cell[:-1]).sum())
    print(res, "distinct corner lines / touches per group of G consecutive samples:",
          {G: round(out[G] / tot, 3) for G in out}, "| cell runs / samples:", round(runs_cell / (3 * N), 3))
