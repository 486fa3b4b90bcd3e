"""Diagnostics: per-step host time of bench.py's two arms (device store without sync / host store with the loss read back),
with the slow steps broken down by phase."""
import sys, time, gc
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
import torch
import bench
from tinynerf_b200 import synthetic
from tinynerf_b200.run import RayStore, TrainConfig, Trainer

dev = torch.device("cuda", 0)
o, d, rgbs, _ = bench.make_scene("blender", bench.N_STORE, bench.SEED)
analytic = synthetic.analytic_grid(128, seed=bench.SEED + 2).to(dev)
amean = analytic.mean().item()
for host in (False, True):
    torch.manual_seed(bench.SEED)
    tr = Trainer(TrainConfig(method="kplanes", scene_type="aabb", batch_size=1024, n_samples=256, seed=bench.SEED),
                 RayStore(o, d, rgbs, dev, host=host, seed=bench.SEED), dev)
    tr.occupancy_grid.grid.copy_(analytic); tr.occupancy_grid.mean = amean
    def pin(t):
        t.occupancy_grid.grid.copy_(analytic); t.occupancy_grid.mean = amean
    tr.post_update = pin
    log = []
    def wrap(obj, name, label):
        f = getattr(obj, name)
        def g(*a, **k):
            t0 = time.perf_counter(); r = f(*a, **k); log.append((label, 1e3 * (time.perf_counter() - t0)))
            return r
        setattr(obj, name, g)
    wrap(tr, "_take_batch", "take"); wrap(tr, "next_batch", "next_batch"); wrap(tr._fused, "forward_backward", "fwd_bwd")
    wrap(tr.optimizer, "step", "adam"); wrap(tr.ray_provider, "count", "count"); wrap(tr.ray_provider, "pack", "pack")
    wrap(tr.store, "next", "store.next"); wrap(tr, "update_occupancy", "update"); wrap(tr, "_prefetch", "prefetch")
    oc = gc.collect
    def tc(*a):
        t0 = time.perf_counter(); r = oc(*a); log.append(("gc", 1e3 * (time.perf_counter() - t0))); return r
    gc.collect = tc
    times = []
    torch.cuda.synchronize()
    for it in range(140):
        log.clear()
        t0 = time.perf_counter()
        info = tr.step()
        t1 = time.perf_counter()
        if host:
            float(info["loss"])
        dt = 1e3 * (time.perf_counter() - t0)
        times.append(dt)
        if it > 3 and dt > 3.5:
            print(f"host={host} step {it}: {dt:.2f} ms (loss read {1e3*(time.perf_counter()-t1):.2f}) | " + " ".join(f"{l}={v:.2f}" for l, v in log if v > 0.3))
    torch.cuda.synchronize()
    ts = sorted(times[4:])
    print(f"host={host}: mean {sum(times[4:])/len(times[4:]):.3f} p50 {ts[len(ts)//2]:.3f} p90 {ts[int(len(ts)*.9)]:.3f} max {ts[-1]:.3f}")
    gc.collect = oc
    del tr
