import sys, copy
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
import torch
from tinynerf_b200 import models
DEV = "cuda"
torch.manual_seed(1)
def run(name, mod, make_in, ref_fn, m):
    gen = torch.Generator().manual_seed(m)
    ins = make_in(m, gen)
    leaf = [t.clone().requires_grad_(True) for t in ins]
    out = mod(*leaf)
    go = torch.randn(out.shape, generator=gen).to(DEV)
    (out * go).sum().backward()
    mod64 = copy.deepcopy(mod).double()
    for p in mod.parameters(): pass
    leaf64 = [t.double().clone().requires_grad_(True) for t in ins]
    out64 = ref_fn(mod64, *leaf64)
    (out64 * go.double()).sum().backward()
    res = [f"out {((out.double()-out64).abs().max()/out64.abs().max()).item():.1e}"]
    for a, b, nm in zip(leaf, leaf64, ["in%d" % i for i in range(len(leaf))]):
        if a.grad is not None:
            res.append(f"{nm} {((a.grad.double()-b.grad).abs().max()/b.grad.abs().max()).item():.1e}")
    for (k, p), (_, q) in zip(mod.named_parameters(), mod64.named_parameters()):
        res.append(f"{k.replace('net.net.','')} {((p.grad.double()-q.grad).abs().max()/q.grad.abs().max().clamp_min(1e-30)).item():.1e}")
        p.grad = None
    print(name, m, " | ".join(res), flush=True)

sig = models.VanillaOpacityDecoder(96).to(DEV)
col = models.VanillaColorDecoder(8, 96, 64, 3).to(DEV)
trunk = models.MLP(36, 128, 5).to(DEV)
mlp2 = models.MLP(64, 64, 1, 3).to(DEV)
for m in (77, 128, 129, 1000, 5000, 40000):
    run("sig  ", sig, lambda m, g: [(torch.randn(m, 96, generator=g) * 0.5).to(DEV)], lambda mm, f: torch.exp(mm.net.net(f) - 1.0), m)
    run("col  ", col, lambda m, g: [(torch.randn(m, 96, generator=g) * 0.5).to(DEV), torch.nn.functional.normalize(torch.randn(m, 3, generator=g), dim=-1).to(DEV)],
        lambda mm, f, d: torch.sigmoid(mm.net.net(torch.cat([mm.pe(d), d, f], -1))), m)
    run("trunk", trunk, lambda m, g: [torch.randn(m, 36, generator=g).to(DEV)], lambda mm, z: mm.net(z), m)
    run("mlp2 ", mlp2, lambda m, g: [torch.randn(m, 64, generator=g).to(DEV)], lambda mm, z: mm.net(z), m)
