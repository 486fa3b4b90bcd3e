"""Diagnostics: where the host time of an occupancy-update iteration goes."""
import sys, time, gc
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
import torch
import bench
from tinynerf_b200 import synthetic
from tinynerf_b200.run import RayStore, TrainConfig, Trainer

dev = torch.device("cuda", 0)
o, d, rgbs, _ = bench.make_scene("blender", 1 << 20, bench.SEED)
tr = Trainer(TrainConfig(method="kplanes", scene_type="aabb", batch_size=1024, n_samples=256, seed=1), RayStore(o, d, rgbs, dev, seed=1), dev)
analytic = synthetic.analytic_grid(128, seed=bench.SEED + 2).to(dev)
tr.occupancy_grid.grid.copy_(analytic); tr.occupancy_grid.mean = analytic.mean().item()
orig_update = tr.update_occupancy
def timed_update():
    torch.cuda.synchronize(); t0 = time.perf_counter()
    orig_update()
    torch.cuda.synchronize(); print(f"   update_occupancy {1e3*(time.perf_counter()-t0):.2f} ms (after draining the queue)")
tr.update_occupancy = timed_update
def wrap(obj, name, label):
    f = getattr(obj, name)
    def g(*a, **k):
        t0 = time.perf_counter(); r = f(*a, **k); dt = 1e3 * (time.perf_counter() - t0)
        if dt > 1.5: print(f"   {label} {dt:.2f} ms")
        return r
    setattr(obj, name, g)
wrap(tr, "_take_batch", "_take_batch"); wrap(tr, "next_batch", "next_batch"); wrap(tr._fused, "forward_backward", "forward_backward")
wrap(tr.optimizer, "step", "optimizer.step"); wrap(tr.ray_provider, "count", "provider.count"); wrap(tr.ray_provider, "pack", "provider.pack")
wrap(tr.store, "next", "store.next")
orig_collect = gc.collect
def timed_collect(*a):
    t0 = time.perf_counter(); r = orig_collect(*a); print(f"   gc.collect {1e3*(time.perf_counter()-t0):.2f} ms -> {r} objects"); return r
gc.collect = timed_collect
for it in range(135):
    t0 = time.perf_counter()
    tr.step()
    dt = 1e3 * (time.perf_counter() - t0)
    if dt > 4 or it % 64 == 0:
        print(f"step {it}: host {dt:.2f} ms")
torch.cuda.synchronize()
print(torch.cuda.memory_summary(abbreviated=True)[:1500])
