import sys, copy
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
import torch
from tinynerf_b200 import _lib, mlp_ops, models
DEV = "cuda"
def rel(a, b): return ((a.double() - b.double()).abs().max() / b.double().abs().max().clamp_min(1e-30)).item()
torch.manual_seed(1)
sig = models.VanillaOpacityDecoder(96).to(DEV)
m = 5000
g = torch.Generator().manual_seed(m)
f0 = (torch.randn(m, 96, generator=g) * 0.5).to(DEV)
go = torch.randn(m, 1, generator=g).to(DEV)
# reference in fp64
s64 = copy.deepcopy(sig).double()
f64 = f0.double().requires_grad_(True)
o64 = torch.exp(s64.net.net(f64) - 1.0)
(o64 * go.double()).sum().backward()
for trial in range(6):
    f = f0.clone().requires_grad_(True)
    out = sig(f)
    (out * go).sum().backward()
    print("trial", trial, "in0", f"{rel(f.grad, f64.grad):.1e}", " ".join(f"{k.replace('net.net.','')} {rel(p.grad, q.grad):.1e}" for (k, p), (_, q) in zip(sig.named_parameters(), s64.named_parameters())), flush=True)
    # locate bad rows of in0
    bad = ((f.grad.double() - f64.grad).abs().max(1).values > 1e-5 * f64.grad.abs().max()).nonzero().flatten()
    print("    bad rows", bad[:8].tolist(), len(bad), "sync now")
    for p in sig.parameters(): p.grad = None
    torch.cuda.synchronize()
# same but with a sync between forward and backward
f = f0.clone().requires_grad_(True)
out = sig(f); torch.cuda.synchronize()
(out * go).sum().backward(); torch.cuda.synchronize()
print("with syncs in0", f"{rel(f.grad, f64.grad):.1e}")
