"""Fused decoder-heads forward at the bench shape for ncu: ncu --set full -k regex:heads_fwd ..."""
import sys
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
sys.path.insert(0, str(Path(__file__).resolve().parents[1] / "tests"))
import torch
from tinynerf_b200 import models
import test_gpu_heads as th
torch.manual_seed(0)
sig = models.VanillaOpacityDecoder(96).to("cuda"); col = models.VanillaColorDecoder(8, 96, 64, 3).to("cuda")
n = 1 << 18
feats = torch.randn(n, 96, device="cuda") * 0.5
dirs = torch.nn.functional.normalize(torch.randn(n, 3, device="cuda"), dim=-1)
for _ in range(3):
    th._run_fused(sig, col, feats, dirs)
print("done")
