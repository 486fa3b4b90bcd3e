"""Run each hot kernel a few times at representative sizes so `ncu -k regex:...` can capture it.
Usage (GPU box): ncu --set full --clock-control none --import-source on -k regex:'weights|kplanes|march|composite' \
                 -c 14 -o gpurun_out/prof python scripts/profile_kernels.py"""
import sys
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
import torch
from tinynerf_b200 import _cuda, core, models, synthetic

dev = "cuda"
which = sys.argv[1] if len(sys.argv) > 1 else "all"
logn = int(sys.argv[2]) if len(sys.argv) > 2 else 24

if which in ("all", "weights"):
    n = 1 << logn
    sig, info, g = synthetic.packed_rays(n, seed=1000 + logn)
    sig, info, g = sig.to(dev), info.to(dev), g.to(dev)
    steps = torch.full_like(sig, 5.196 / 256)
    for _ in range(2):
        w = _cuda.weights_fwd(sig, steps, info, 1e-4, _cuda.TRUSTED_PARTITION)
        gs = _cuda.weights_bwd(sig, steps, info, w, g, _cuda.TRUSTED_PARTITION)
    torch.cuda.synchronize()

if which in ("all", "kplanes"):
    torch.manual_seed(0)
    field = models.KPlanesFeatureField(32).to(dev)
    aabb = torch.tensor([[-1.5, -1.5, -1.5], [1.5, 1.5, 1.5]], device=dev)
    marcher = core.RayMarcherAABB(aabb, 256, 0.1)
    og = core.OccupancyGrid(128, marcher.step_size, 0.01, synthetic.DECAY).to(dev)
    og.grid.copy_(synthetic.analytic_grid(128, seed=1236))
    og.mean = og.grid.mean().item()
    prov = core.RayProvider(og, core.ContractionAABB(aabb), marcher)
    o, d = synthetic.blender_rays(9500, seed=2)
    for _ in range(2):
        packed, info = prov(o.to(dev), d.to(dev), training=True)
        feats = field(packed[:, :3])
        feats.backward(torch.randn_like(feats))
        w = torch.rand(packed.size(0), device=dev, requires_grad=True)
        rgb = torch.rand(packed.size(0), 3, device=dev, requires_grad=True)
        out = core.Composite.apply(w, rgb, info, [1.0, 1.0, 1.0])
        out.sum().backward()
    print("packed", packed.shape)
    torch.cuda.synchronize()
