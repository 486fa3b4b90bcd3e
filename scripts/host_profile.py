"""Where the HOST time of a training iteration goes (cProfile over N steps of the bench workload).
`python scripts/host_profile.py [kplanes|cobafa|vanilla]`"""
import cProfile
import pstats
import sys
import time
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
import torch
import bench
from tinynerf_b200 import synthetic
from tinynerf_b200.run import RayStore, TrainConfig, Trainer

dev = torch.device("cuda", 0)
o, d, rgbs, _ = bench.make_scene("blender", 1 << 20, bench.SEED)
method = sys.argv[1] if len(sys.argv) > 1 else "kplanes"
cfg = TrainConfig(method=method, scene_type="aabb", batch_size=1024, n_samples=256, seed=1)
tr = Trainer(cfg, RayStore(o, d, rgbs, dev, seed=1), dev)
tr.occupancy_grid.grid.copy_(synthetic.analytic_grid(128, seed=bench.SEED + 2).to(dev))
tr.occupancy_grid.mean = tr.occupancy_grid.grid.mean().item()
tr.train_step = 1
for _ in range(5):
    tr.step()
torch.cuda.synchronize()
n = 40
t0 = time.perf_counter()
pr = cProfile.Profile()
pr.enable()
for _ in range(n):
    tr.step()
pr.disable()
t1 = time.perf_counter()
torch.cuda.synchronize()
t2 = time.perf_counter()
print(f"host {1e3 * (t1 - t0) / n:.3f} ms/step (under cProfile), +sync {1e3 * (t2 - t1):.3f} ms")
st = pstats.Stats(pr)
st.sort_stats("cumulative").print_stats(45)
st.sort_stats("tottime").print_stats(25)
