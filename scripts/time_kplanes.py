"""K-Planes lookup fwd/bwd device time at the bench shape (graph replay over rotating inputs)."""
import sys
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
import ctypes as C
import torch
from tinynerf_b200 import _lib, core, models, synthetic

dev = "cuda"
torch.manual_seed(0)
field = models.KPlanesFeatureField(32).to(dev)
aabb = torch.tensor([[-1.5, -1.5, -1.5], [1.5, 1.5, 1.5]], device=dev)
marcher = core.RayMarcherAABB(aabb, 256, 0.1)
og = core.OccupancyGrid(128, marcher.step_size, 0.01, synthetic.DECAY).to(dev)
og.grid.copy_(synthetic.analytic_grid(128, seed=1236))
og.mean = og.grid.mean().item()
prov = core.RayProvider(og, core.ContractionAABB(aabb), marcher)
packs = []
for i in range(4):
    o, d = synthetic.blender_rays(9500, seed=2 + i)
    packed, info = prov(o.to(dev), d.to(dev), training=True)
    packs.append(packed)
n = min(p.size(0) for p in packs)
print("samples per launch", n)
planes = field._plane_params()
stor = [models._channels_last_storage(p) for p in planes]
ptrs = (C.c_void_p * 9)(*[t.data_ptr() for t in stor])
grads = [torch.zeros_like(t) for t in stor]
gptrs = (C.c_void_p * 9)(*[t.data_ptr() for t in grads])
res = (C.c_int32 * 3)(128, 256, 512)
out = torch.empty(n, 96, device=dev)
go = torch.randn(n, 96, device=dev)
flush = torch.empty(64 << 20, device=dev)  # 256 MB


def timeit(fn, reps=8):
    fn(0); torch.cuda.synchronize()
    ts = []
    for k in range(reps):
        flush.zero_()  # evict the planes from L2, as the optimiser pass does between launches in a training step
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record(); fn(k % 4); e.record(); torch.cuda.synchronize()
        ts.append(s.elapsed_time(e))
    return sorted(ts)[len(ts) // 2] * 1e3


st = _lib.stream_ptr()
fwd = lambda i: _lib.call("tnf_kplanes_fwd", ptrs, res, 3, 32, packs[i].data_ptr(), 7, n, out.data_ptr(), st)
bwd = lambda i: _lib.call("tnf_kplanes_bwd", ptrs, gptrs, res, 3, 32, packs[i].data_ptr(), 7, n, go.data_ptr(), st)
tf, tb = timeit(fwd), timeit(bwd)
P = 132120576
print(f"kplanes fwd {tf:.1f} us  alg {(n * 396 + P) / tf / 1e3:.0f} GB/s | bwd {tb:.1f} us alg {(n * 396 + 2 * P) / tb / 1e3:.0f} GB/s")

# per-scale breakdown (n_scales = 1 calls)
out1 = torch.empty(n, 32, device=dev)
go1 = torch.randn(n, 32, device=dev)
for s, r in enumerate((128, 256, 512)):
    p1 = (C.c_void_p * 3)(*[t.data_ptr() for t in stor[3 * s:3 * s + 3]])
    g1 = (C.c_void_p * 3)(*[t.data_ptr() for t in grads[3 * s:3 * s + 3]])
    r1 = (C.c_int32 * 1)(r)
    f1 = lambda i: _lib.call("tnf_kplanes_fwd", p1, r1, 1, 32, packs[i].data_ptr(), 7, n, out1.data_ptr(), st)
    b1 = lambda i: _lib.call("tnf_kplanes_bwd", p1, g1, r1, 1, 32, packs[i].data_ptr(), 7, n, go1.data_ptr(), st)
    print(f"  scale {r}: fwd {timeit(f1):.1f} us | bwd {timeit(b1):.1f} us")
# without the L2 flush between launches (planes possibly L2-resident)
def timeit_warm(fn, reps=8):
    fn(0); torch.cuda.synchronize()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    for k in range(reps):
        fn(k % 4)
    e.record(); torch.cuda.synchronize()
    return s.elapsed_time(e) / reps * 1e3
print(f"  back-to-back (no flush): fwd {timeit_warm(fwd):.1f} us | bwd {timeit_warm(bwd):.1f} us")

# ---- round 2: where does the scatter's time go? (experimental variants, tnf_kplanes_bwd_ex) ----
for mode, label in ((1, "gather+blend only"), (2, "red.v4 scatter only"), (3, "gather + TMA bulk-reduce scatter")):
    bx = lambda i: _lib.call("tnf_kplanes_bwd_ex", ptrs, gptrs, res, 3, 32, packs[i].data_ptr(), 7, n, go.data_ptr(), mode, st)
    print(f"  bwd variant {mode} ({label}): flushed {timeit(bx):.1f} us | back-to-back {timeit_warm(bx):.1f} us")
# correctness of the TMA variant against the red.v4 kernel
for g_ in grads: g_.zero_()
bwd(0); torch.cuda.synchronize()
ref = [g_.clone() for g_ in grads]
for g_ in grads: g_.zero_()
_lib.call("tnf_kplanes_bwd_ex", ptrs, gptrs, res, 3, 32, packs[0].data_ptr(), 7, n, go.data_ptr(), 3, st); torch.cuda.synchronize()
print("  TMA scatter max rel err vs red.v4:", max(float((a - b).abs().max() / b.abs().max()) for a, b in zip(grads, ref)))

# ---- occupancy variants (tnf_set_variant 2: blocks per SM the kernel is compiled for) ----
for k in (0, 1, 5, 6, 8):   # 0 = the default builds (forward 6, backward 4), 1 = uncapped
    _lib.load().tnf_set_variant(2, k)
    print(f"  occupancy variant {k}: fwd {timeit(fwd):.1f} us | bwd {timeit(bwd):.1f} us")
_lib.load().tnf_set_variant(2, 0)
