"""Diagnostics: wait cycles per warp role of heads_fwd_kernel (CTA 0)."""
import sys, ctypes as C
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
sys.path.insert(0, str(Path(__file__).resolve().parents[1] / "tests"))
import torch
from tinynerf_b200 import _lib, models
import test_gpu_heads as th
lib = _lib.load()
buf = torch.zeros(32, dtype=torch.int64, device="cuda")
lib.tnf_debug_heads_timing.argtypes = [C.c_void_p]
assert lib.tnf_debug_heads_timing(buf.data_ptr()) == 0
torch.manual_seed(0)
sig = models.VanillaOpacityDecoder(96).to("cuda"); col = models.VanillaColorDecoder(8, 96, 64, 3).to("cuda")
n = 1 << 18
feats = torch.randn(n, 96, device="cuda") * 0.5
dirs = torch.nn.functional.normalize(torch.randn(n, 3, device="cuda"), dim=-1)
for _ in range(2):
    th._run_fused(sig, col, feats, dirs)
v = buf.tolist()
print(f"loader: wait_ahempty {v[0]} wait_wempty {v[1]} wait_cp {v[2]} wait_alempty {v[3]} lo_pass {v[4]} total {v[5]} units {v[6]}")
print(f"mma: wait_dsempty {v[8]} wait_afull {v[9]} wait_wfull {v[10]} wait_act {v[11]} issue_l0 {v[12]} issue_hidden {v[13]} total {v[14]}")
print(f"colour epi: wait {v[16]} total {v[17]} | sigma epi: wait {v[20]} total {v[21]}")
print(f"colour epi pieces: tmem_ld {v[24]} bias/relu {v[25]} store_rows {v[26]} tmem_st {v[27]}")
print(f"loader: cp.async atom issue {v[7]} fence+arrive {v[15]}")
