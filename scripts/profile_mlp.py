"""MLP kernels at the bench shape for ncu: ncu --set full -k regex:'linear|wgrad|head_bwd|color_input' ..."""
import sys
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
import torch
from tinynerf_b200 import models
dev = "cuda"
torch.manual_seed(0)
sig = models.VanillaOpacityDecoder(96).to(dev)
col = models.VanillaColorDecoder(8, 96, 64, 3).to(dev)
m = 1 << 18
f = (torch.randn(m, 96, device=dev) * 0.5).requires_grad_(True)
d = torch.nn.functional.normalize(torch.randn(m, 3, device=dev), dim=-1)
for _ in range(2):
    s = sig(f); c = col(f, d)
    (s.sum() + c.sum()).backward()
torch.cuda.synchronize()
print("done")
