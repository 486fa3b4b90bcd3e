# scratch timing helper for early GPU runs (not part of the product)
import torch, time, sys
sys.path.insert(0, str(__import__('pathlib').Path(__file__).resolve().parents[1]))
from tinynerf_b200 import _cuda, synthetic, core, models
import oracle
dev = 'cuda'
ref = oracle.load_ref_cuda()
def timeit(f, n=20, w=3):
    for _ in range(w): f()
    torch.cuda.synchronize()
    s = torch.cuda.Event(enable_timing=True); e = torch.cuda.Event(enable_timing=True)
    s.record()
    for _ in range(n): f()
    e.record(); torch.cuda.synchronize()
    return s.elapsed_time(e) / n
for logn in (18, 22, 26):
    n = 1 << logn
    sig, info, g = synthetic.packed_rays(n, seed=1000 + logn)
    sig, info, g = sig.to(dev), info.to(dev), g.to(dev)
    steps = torch.full_like(sig, 5.196 / 256)
    R = info.size(0)
    for flags, name in ((0, 'untrusted'), (1, 'trusted'), (3, 'trusted,noexact')):
        tf = timeit(lambda: _cuda.weights_fwd(sig, steps, info, 1e-4, flags))
        w = _cuda.weights_fwd(sig, steps, info, 1e-4, flags)
        tb = timeit(lambda: _cuda.weights_bwd(sig, steps, info, w, g, flags))
        bf = (12 * n + 8 * R) / tf / 1e6; bb = (20 * n + 8 * R) / tb / 1e6
        print(f"N=2^{logn} R={R} {name}: fwd {tf*1e3:.1f} us {bf:.0f} GB/s | bwd {tb*1e3:.1f} us {bb:.0f} GB/s", flush=True)
    if R <= (1 << 20):
        tf = timeit(lambda: ref.compute_weights_fwd(sig, steps, info, 1e-4))
        tb = timeit(lambda: ref.compute_weights_bwd(sig, steps, info, w, g))
        print(f"   reference kernel: fwd {tf*1e3:.1f} us | bwd {tb*1e3:.1f} us", flush=True)
# kplanes
torch.manual_seed(0)
field = models.KPlanesFeatureField(32).to(dev)
n = 1 << 18
x = (torch.rand(n, 3, device=dev) * 2 - 1)
t = timeit(lambda: field(x))
print(f"kplanes fwd N=2^18: {t*1e3:.1f} us -> {(n*396+132120576)/t/1e6:.0f} GB/s alg", flush=True)
go = torch.randn(n, 96, device=dev)
def fb():
    field.zero_grad(set_to_none=True)
    (field(x)).backward(go)
t2 = timeit(fb)
print(f"kplanes fwd+bwd(+zeros): {t2*1e3:.1f} us", flush=True)
