"""Generate tests/golden/*.npz by running the REAL reference (imported from /root/reference) on CPU.

Run in the build container only (the GPU box has no /root/reference):
    python tests/golden/make_golden.py

The reference JIT-builds its CUDA op at import (src/core.py:7) and that op rejects CPU tensors
(src/cuda.cu:62), so `torch.utils.cpp_extension.load` is stubbed to return the C oracle's weights
functions; every other line executed below is the reference's own code.  Inputs are seeded and stored
next to the outputs, so the fixtures are self-contained.
"""
import os
import sys
import types
from pathlib import Path

import numpy as np
import torch

ROOT = Path(__file__).resolve().parents[2]
REF = Path(os.environ.get("TNF_REFERENCE_ROOT", "/root/reference"))
sys.path.insert(0, str(ROOT))
sys.dont_write_bytecode = True

from oracle import c as orc  # noqa: E402

stub = types.SimpleNamespace(compute_weights_fwd=orc.weights_fwd, compute_weights_bwd=orc.weights_bwd)
import torch.utils.cpp_extension as cpp  # noqa: E402

cpp.load = lambda *a, **k: stub
sys.path.insert(0, str(REF))
os.chdir(REF)
from src import core as rcore  # noqa: E402
from src import models as rmodels  # noqa: E402

OUT = Path(__file__).resolve().parent
DECAY = 0.01 ** (1 / 16)


def save(name, **arrays):
    np.savez_compressed(OUT / f"{name}.npz", **{k: (v.detach().cpu().numpy() if isinstance(v, torch.Tensor) else np.asarray(v))
                                                for k, v in arrays.items()})
    print("wrote", name, sorted(arrays))


def random_grid(res, seed):
    g = torch.Generator().manual_seed(seed)
    k = torch.randint(0, 25, (res, res, res), generator=g)
    grid = torch.tensor(DECAY, dtype=torch.float32) ** k.float()
    grid[k == 0] = 1.0
    return grid


def camera_rays(n, seed, radius=4.0):
    g = torch.Generator().manual_seed(seed)
    o = torch.nn.functional.normalize(torch.randn(n, 3, generator=g), dim=-1) * radius
    target = (torch.rand(n, 3, generator=g) - 0.5) * 2.0
    d = torch.nn.functional.normalize(target - o, dim=-1)
    d[0, 1] = 0.0  # exercise the d == 0 branch (src/core.py:78)
    return o, d


def provider_case(name, scene, seed):
    R, S, res = 48, 64, 32
    o, d = camera_rays(R, seed)
    grid_vals = random_grid(res, seed + 1)
    aabb = torch.tensor([[-1.5, -1.5, -1.5], [1.5, 1.5, 1.5]])
    if scene == "aabb":
        marcher = rcore.RayMarcherAABB(aabb, S, 0.1)
        contraction = rcore.ContractionAABB(aabb)
    else:
        marcher = rcore.RayMarcherUnbounded(S, 0.1, 1e5, uniform_range=1.7)
        contraction = rcore.ContractionMip360(order=float("inf"))
    og = rcore.OccupancyGrid(res, marcher.step_size, 0.01, DECAY)
    og.grid.copy_(grid_vals)
    og.mean = og.grid.mean().item()
    rp = rcore.RayProvider(og, contraction, marcher)
    packed_eval, info_eval = rp(o, d, training=False)
    torch.manual_seed(seed + 2)
    packed_tr, info_tr = rp(o, d, training=True)
    torch.manual_seed(seed + 2)
    noise = torch.rand(R, S)  # same generator consumption as rand_like(t_values) (src/core.py:173)
    t_values, step_sizes = marcher(o, d)
    save(name, rays_o=o, rays_d=d, grid=grid_vals, aabb=aabb, threshold=np.float64(og.threshold),
         step_size=np.float32(float(marcher.step_size)), noise=noise, t_values=t_values.contiguous(),
         step_sizes=step_sizes.contiguous(), packed_eval=packed_eval, info_eval=info_eval,
         packed_train=packed_tr, info_train=info_tr)


def occupancy_update_case():
    res = 16
    og = rcore.OccupancyGrid(res, 0.02, 0.01, DECAY)
    og.grid.copy_(random_grid(res, 5))
    og.mean = og.grid.mean().item()
    before = og.grid.clone()
    thr_before = og.threshold

    def sigma_fn(x):
        return 60.0 * torch.exp(-6.0 * (x ** 2).sum(-1, keepdim=True))

    torch.manual_seed(11)
    og.update(sigma_fn)
    torch.manual_seed(11)
    noise = torch.stack([torch.rand(res, res, 3) for _ in range(res)])
    save("occ_update", grid_before=before, grid_after=og.grid, noise=noise, mean_after=np.float64(og.mean),
         threshold_before=np.float64(thr_before), step_size=np.float32(0.02), decay=np.float64(DECAY))


def kplanes_case():
    torch.manual_seed(21)
    field = rmodels.KPlanesFeatureField(32)
    g = torch.Generator().manual_seed(22)
    x = torch.rand(96, 3, generator=g) * 2.2 - 1.1  # a few points outside [-1,1]: zero padding
    x[0] = torch.tensor([-1.0, 1.0, 0.0])
    feats = field(x)
    go = torch.randn(feats.shape, generator=g)
    (feats * go).sum().backward()
    probes = {}
    for s, scale in enumerate(field.planes):
        for p, plane in enumerate(scale):
            gr = plane.plane.grad
            nz = gr.abs().reshape(-1).topk(64).indices
            probes[f"gidx_{s}_{p}"] = nz
            probes[f"gval_{s}_{p}"] = gr.reshape(-1)[nz]
            probes[f"gsum_{s}_{p}"] = gr.double().sum()
            probes[f"gabs_{s}_{p}"] = gr.double().abs().sum()
    save("kplanes", x=x, features=feats, grad_out=go, tv=field.loss_tv(), l1=field.loss_l1(),
         param_checksum=sum(p.double().sum() for p in field.parameters()), **probes)


def cobafa_case():
    torch.manual_seed(31)
    field = rmodels.CobafaFeatureField(basis_res=[8, 11, 14], coef_res=9, freqs=torch.linspace(2.0, 8.0, 3).tolist(),
                                       channels=[8, 4, 2], mlp_hidden_dim=32)
    field.eval()
    g = torch.Generator().manual_seed(32)
    x = torch.rand(80, 3, generator=g) * 2.2 - 1.1
    coefs = field.coef_grid(x)
    ys = [basis(enc(x)) * coefs[:, [i]] for i, (enc, basis) in enumerate(zip(field.encoders, field.basis_grids))]
    lookup = torch.cat(ys, -1)
    go = torch.randn(lookup.shape, generator=g)
    (lookup * go).sum().backward()
    grads = {f"gbasis_{i}": b.grid.grad for i, b in enumerate(field.basis_grids)}
    grads["gcoef"] = field.coef_grid.grid.grad
    out = field(x)
    save("cobafa", x=x, lookup=lookup, grad_out=go, forward_eval=out, **grads)


def heads_case():
    torch.manual_seed(41)
    sig = rmodels.VanillaOpacityDecoder(96)
    col = rmodels.VanillaColorDecoder(8, 96, 64, 3)
    g = torch.Generator().manual_seed(42)
    feats = torch.randn(64, 96, generator=g) * 0.5
    dirs = torch.nn.functional.normalize(torch.randn(64, 3, generator=g), dim=-1)
    pe = rmodels.PositionalEncoding(8)(dirs)
    save("heads", feats=feats, dirs=dirs, sigma=sig(feats), rgb=col(feats, dirs), pe=pe)


def renderer_case():
    """Full provider -> renderer pipeline (vanilla model, AABB) on CPU, weights via the C oracle."""
    torch.manual_seed(51)
    fm = rmodels.VanillaFeatureMLP(4, 32, 1)
    sd = rmodels.VanillaOpacityDecoder(32)
    cd = rmodels.VanillaColorDecoder(4, 32, 32, 1)
    aabb = torch.tensor([[-1.5, -1.5, -1.5], [1.5, 1.5, 1.5]])
    marcher = rcore.RayMarcherAABB(aabb, 48, 0.1)
    og = rcore.OccupancyGrid(16, marcher.step_size, 0.01, DECAY)
    og.grid.copy_(random_grid(16, 52))
    og.mean = og.grid.mean().item()
    rp = rcore.RayProvider(og, rcore.ContractionAABB(aabb), marcher)
    bg = torch.tensor([1.0, 1.0, 1.0])
    renderer = rcore.NerfRenderer(fm, sd, cd, bg_color=bg)
    with torch.no_grad():  # make the volume opaque enough for early termination to trigger
        sd.net.net[-1].bias += 5.0
    o, d = camera_rays(40, 53)
    packed, info = rp(o, d, training=False)
    out = renderer(packed, info)
    loss = ((out - 0.25) ** 2).mean()
    loss.backward()
    save("renderer", rays_o=o, rays_d=d, grid=og.grid, packed=packed, info=info, rendered=out,
         loss=loss, grad_sigma_w=sd.net.net[0].weight.grad, grad_feat_w=fm.net.net[0].weight.grad)


if __name__ == "__main__":
    provider_case("provider_aabb", "aabb", 100)
    provider_case("provider_unbounded", "unbounded", 200)
    occupancy_update_case()
    kplanes_case()
    cobafa_case()
    heads_case()
    renderer_case()
