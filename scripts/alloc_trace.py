"""Which step of the bench workload makes PyTorch's caching allocator call cudaMalloc (new segment), and for what size?"""
import sys
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
import torch
import bench
from tinynerf_b200 import synthetic
from tinynerf_b200.run import RayStore, TrainConfig, Trainer

dev = torch.device("cuda", 0)
o, d, rgbs, _ = bench.make_scene("blender", bench.N_STORE, bench.SEED)
torch.manual_seed(bench.SEED)
cfg = TrainConfig(method="kplanes", scene_type="aabb", batch_size=bench.BATCH, n_samples=bench.N_SAMPLES, seed=bench.SEED)
tr = Trainer(cfg, RayStore(o, d, rgbs, dev, seed=bench.SEED), dev)
analytic = synthetic.analytic_grid(128, seed=bench.SEED + 2).to(dev)
tr.occupancy_grid.grid.copy_(analytic); tr.occupancy_grid.mean = analytic.mean().item()
tr.post_update = lambda t: (t.occupancy_grid.grid.copy_(analytic), setattr(t.occupancy_grid, "mean", analytic.mean().item()))
stat = lambda k: torch.cuda.memory_stats(dev).get(k, 0)
seg = stat("segment.all.allocated")
import time
slow = []
for step in range(900):
    t0 = time.perf_counter()
    info = tr.step()
    dt = (time.perf_counter() - t0) * 1e3
    if dt > 2.5:
        slow.append((step, round(dt, 2)))
    s2 = stat("segment.all.allocated")
    if s2 != seg:
        print(f"step {step}: +{s2 - seg} segment(s); n_samples {info['n_samples']} n_rays {info['n_rays']} chunks_guess {tr._chunks_guess:.2f} "
              f"reserved {stat('reserved_bytes.all.current') >> 20} MB active {stat('active_bytes.all.current') >> 20} MB "
              f"largest free? inactive_split {stat('inactive_split_bytes.all.current') >> 20} MB")
        seg = s2
torch.cuda.synchronize()
print("done; segments", seg)
print("host steps slower than 2.5 ms:", slow)
print(torch.cuda.memory_summary(dev, abbreviated=True)[:1500])
