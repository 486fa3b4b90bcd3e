"""Diagnostics: what bounds the device-resident arm of bench.py?  Same trainer, 64 event-timed steps, under a few knobs;
per-step GPU intervals from end-of-step events on the main stream."""
import os, sys, time
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
import torch
from tinynerf_b200 import _lib
import bench
from tinynerf_b200 import synthetic
from tinynerf_b200.run import RayStore, TrainConfig, Trainer

dev = torch.device("cuda", 0)
o, d, rgbs, _ = bench.make_scene("blender", bench.N_STORE, bench.SEED)
analytic = synthetic.analytic_grid(128, seed=bench.SEED + 2).to(dev)
amean = analytic.mean().item()

def run(label, env=None, **kw):
    for k, v in (env or {}).items():
        os.environ[k] = v
    torch.manual_seed(bench.SEED)
    cfg = TrainConfig(method="kplanes", scene_type="aabb", batch_size=1024, n_samples=256, seed=bench.SEED, **kw)
    tr = Trainer(cfg, RayStore(o, d, rgbs, dev, seed=bench.SEED), dev)
    tr.occupancy_grid.grid.copy_(analytic); tr.occupancy_grid.mean = amean
    def pin(t):
        t.occupancy_grid.grid.copy_(analytic); t.occupancy_grid.mean = amean
    tr.post_update = pin
    for _ in range(3):
        tr.step()
    torch.cuda.synchronize()
    evs = [torch.cuda.Event(enable_timing=True) for _ in range(65)]
    evs[0].record()
    n = 0
    t0 = time.perf_counter()
    for i in range(64):
        n += tr.step()["n_samples"]
        evs[i + 1].record()
    host = time.perf_counter() - t0
    torch.cuda.synchronize()
    ms = evs[0].elapsed_time(evs[-1])
    per = sorted(evs[i].elapsed_time(evs[i + 1]) for i in range(64))
    print(f"{label:28s}: {n / ms / 1e3:7.1f} M/s  {ms / 64:.4f} ms/step | gpu step p10 {per[6]:.3f} p50 {per[32]:.3f} p90 {per[57]:.3f} max {per[-1]:.3f} "
          f"| host {host * 1e3 / 64:.3f} ms/step", flush=True)
    for k in (env or {}):
        os.environ.pop(k, None)
    del tr
    torch.cuda.empty_cache()

run("default")
run("default again")
_lib.load().tnf_set_variant(0, 1)
run("wgrad SS (slower)")
_lib.load().tnf_set_variant(0, 0)
run("no prefetch", prefetch=False)
run("inflight unbounded", max_inflight_steps=0)
run("inflight 2", max_inflight_steps=2)
run("inflight 6", max_inflight_steps=6)
run("prefetch depth 1", prefetch_depth=1)
run("prefetch depth 4", prefetch_depth=4)
