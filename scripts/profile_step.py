"""A few steady-state training iterations of the bench workload (K-Planes AABB, ~2^18 packed samples) for ncu:
    ncu --set full --clock-control none -k regex:tnf --launch-skip 90 --launch-count 36 -o /tmp/prof_step python scripts/profile_step.py
    python scripts/summarize_ncu.py full /tmp/prof_step.ncu-rep > profiles/rNN_ncu_full.md
(skip = the first two iterations incl. the occupancy update of iteration 0; one iteration is ~27 launches of ours)."""
import sys
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
import torch
import bench
from tinynerf_b200 import synthetic
from tinynerf_b200.run import RayStore, TrainConfig, Trainer

dev = torch.device("cuda", 0)
o, d, rgbs, _ = bench.make_scene("blender", 1 << 18, bench.SEED)
torch.manual_seed(bench.SEED)
cfg = TrainConfig(method="kplanes", scene_type="aabb", batch_size=bench.BATCH, n_samples=bench.N_SAMPLES, seed=bench.SEED, prefetch=False)
tr = Trainer(cfg, RayStore(o, d, rgbs, dev, seed=bench.SEED), dev)
analytic = synthetic.analytic_grid(128, seed=bench.SEED + 2).to(dev)
tr.occupancy_grid.grid.copy_(analytic)
tr.occupancy_grid.mean = analytic.mean().item()
tr.post_update = lambda t: (t.occupancy_grid.grid.copy_(analytic), setattr(t.occupancy_grid, "mean", analytic.mean().item()))
for _ in range(int(sys.argv[1]) if len(sys.argv) > 1 else 4):
    info = tr.step()
torch.cuda.synchronize()
print("done", info["n_samples"])
