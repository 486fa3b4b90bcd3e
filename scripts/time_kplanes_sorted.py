"""Sorted plane scatter (tnf_kplanes_sort + tnf_kplanes_bwd_sorted) against the direct scatter at the bench shape:
sort, phase 1 (gather + finest scale direct + coarse rows stored), phase 2 (run-merging scatter of the coarse scales)."""
import sys
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
import ctypes as C
import torch
from tinynerf_b200 import _lib, core, models, synthetic

dev = "cuda"


def kplanes_case(n_rays, seed=0):
    """BASELINE config 2 shape from the product's own modules: K-Planes + vanilla heads, AABB +-1.5, analytic 128^3 grid."""
    torch.manual_seed(seed)
    field = models.KPlanesFeatureField(32)
    sd, cd = models.VanillaOpacityDecoder(96), models.VanillaColorDecoder(8, 96, 64, 3)
    renderer = core.NerfRenderer(field, sd, cd, bg_color=torch.ones(3)).to(dev)
    aabb = torch.tensor([[-1.5, -1.5, -1.5], [1.5, 1.5, 1.5]], device=dev)
    marcher = core.RayMarcherAABB(aabb, 256, 0.1)
    og = core.OccupancyGrid(128, marcher.step_size, 0.01, synthetic.DECAY).to(dev)
    og.grid.copy_(synthetic.analytic_grid(128, seed=1))
    og.mean = og.grid.mean().item()
    prov = core.RayProvider(og, core.ContractionAABB(aabb), marcher)
    o, d = synthetic.blender_rays(n_rays, seed=2)
    return renderer, prov, og, aabb, o.to(dev), d.to(dev)


renderer, prov, og, aabb, o, d = kplanes_case(9600)
torch.manual_seed(5)
packed, info = prov(o, d, training=True)
n = packed.size(0)
stor = [models._channels_last_storage(p) for p in renderer.feature_module._plane_params()]
ptrs = (C.c_void_p * 9)(*[t.data_ptr() for t in stor])
grads = [torch.zeros_like(t) for t in stor]
gp = (C.c_void_p * 9)(*[t.data_ptr() for t in grads])
res = (C.c_int32 * 3)(128, 256, 512)
go = torch.randn(n, 96, device=dev)
lib = _lib.load()
st = _lib.stream_ptr()
flush = torch.empty(64 << 20, device=dev)


def timeit(fn, reps=7):
    fn(); torch.cuda.synchronize()
    ts = []
    for _ in range(reps):
        flush.zero_()
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record(); fn(); e.record(); torch.cuda.synchronize()
        ts.append(s.elapsed_time(e) * 1e3)
    return sorted(ts)[len(ts) // 2]


print("samples", n)
print(f"direct scatter           : {timeit(lambda: _lib.call('tnf_kplanes_bwd', ptrs, gp, res, 3, 32, packed.data_ptr(), 7, n, go.data_ptr(), st)):.1f} us")
for sort_res, n_sorted in ((256, 2), (128, 1), (512, 3)):
    pos = torch.empty(3, n, dtype=torch.int32, device=dev)
    uv = torch.empty(3, n, 2, device=dev)
    scratch = torch.empty(int(lib.tnf_kplanes_sort_scratch_ints(sort_res, n)), dtype=torch.int32, device=dev)
    rows = torch.empty(3 * n_sorted * n * 32, device=dev)
    sort = lambda: _lib.call("tnf_kplanes_sort", packed.data_ptr(), 7, n, sort_res, scratch.data_ptr(), pos.data_ptr(), uv.data_ptr(), st)
    ph = lambda p: _lib.call("tnf_kplanes_bwd_sorted", ptrs, gp, res, 3, 32, packed.data_ptr(), 7, n, go.data_ptr(), n_sorted, pos.data_ptr(),
                             uv.data_ptr(), rows.data_ptr(), p, st)
    print(f"sorted scales {n_sorted} (sort at {sort_res}): sort {timeit(sort):.1f} us | phase 1 {timeit(lambda: ph(1)):.1f} us | "
          f"phase 2 {timeit(lambda: ph(2)):.1f} us | both {timeit(lambda: ph(0)):.1f} us")
