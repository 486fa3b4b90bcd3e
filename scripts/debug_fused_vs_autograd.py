"""Diagnostics: per-parameter gradient differences between the fused iteration and the module/autograd path."""
import sys
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
sys.path.insert(0, str(Path(__file__).resolve().parents[1] / "tests"))
import torch
from test_gpu_fused import _trainer

res = {}
for fused in (False, True):
    tr = _trainer(fused)
    tr.optimizer.step = lambda *a, **k: None
    torch.manual_seed(10)
    info = tr.step()
    res[fused] = ({k: p.grad.clone() for k, p in tr.renderer.named_parameters()}, float(info["loss"]), info["n_samples"])
print("n", res[True][2], res[False][2], "loss", res[True][1], res[False][1])
for k, b in res[False][0].items():
    a = res[True][0][k]
    scale = b.abs().max().clamp_min(1e-12)
    err = (a - b).abs()
    print(f"{k:45s} numel {a.numel():9d} scale {float(scale):.3e} max err/scale {float(err.max() / scale):.3e} "
          f"frac>1e-5 {float((err > 1e-5 * scale).float().mean()):.2e} frac>1e-4 {float((err > 1e-4 * scale).float().mean()):.2e}")
