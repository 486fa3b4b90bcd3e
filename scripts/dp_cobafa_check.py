"""Data-parallel Cobafa (torchrun, G ranks): the fused iteration (fused_cobafa.py: ray count + ONE flat-gradient all-reduce +
replicated Adam) against the module/autograd path (per-step cat + all-reduce + copy) on the same ray shards and seeds.
Checks: the replicas stay bit-identical across ranks (parameters after K steps), and both paths produce the same losses
(1e-4 relative: float-atomics order differs) and parameters (fraction of entries off by > 2e-4 of the tensor max <= 3 %).
    python -m torch.distributed.run --nproc-per-node 2 --master-addr 127.0.0.1 scripts/dp_cobafa_check.py"""
import json
import os
import sys
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
import torch
import torch.distributed as dist
from tinynerf_b200 import synthetic
from tinynerf_b200.run import RayStore, TrainConfig, Trainer

rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
torch.cuda.set_device(int(os.environ.get("LOCAL_RANK", rank)))
dev = torch.device("cuda", torch.cuda.current_device())
dist.init_process_group("nccl", device_id=dev)
o, d = synthetic.blender_rays(1 << 16, seed=3)
rgb = torch.rand(1 << 16, 3, generator=torch.Generator().manual_seed(4))
K = 4
res = {}
for fused in (True, False):
    cfg = TrainConfig(method="cobafa", scene_type="aabb", batch_size=256, n_samples=128, prefetch=False, fused_step=fused)
    torch.manual_seed(9)
    tr = Trainer(cfg, RayStore(o, d, rgb, dev, seed=1, rank=rank, world=world), dev, rank=rank, world=world)
    assert (tr._fused_cobafa is not None) == fused
    tr.occupancy_grid.grid.copy_(synthetic.analytic_grid(128, seed=5).to(dev))
    tr.occupancy_grid.mean = tr.occupancy_grid.grid.mean().item()
    tr.train_step = 1
    losses = []
    for it in range(K):
        torch.manual_seed(100 + it + 1000 * rank)
        out = tr.step()
        l = out["loss"].detach().clone().float()
        dist.all_reduce(l)          # every rank reports its share of the union loss
        losses.append(float(l))
    params = {k: p.detach().clone() for k, p in tr.renderer.named_parameters()}
    same = True
    for k, p in params.items():
        flat = torch.as_strided(p, (p.numel(),), (1,), p.storage_offset()) if not p.is_contiguous() else p.view(-1)
        hi, lo = flat.clone(), flat.clone()
        dist.all_reduce(hi, op=dist.ReduceOp.MAX)
        dist.all_reduce(lo, op=dist.ReduceOp.MIN)
        same &= bool(torch.equal(hi, lo))
    res[fused] = (losses, params, same)
    tr.close()
rep = {"world": world, "steps": K, "replicas_bit_identical": {"fused": res[True][2], "modules": res[False][2]},
       "union_loss_fused": res[True][0], "union_loss_modules": res[False][0]}
worst_frac, worst_key = 0.0, None
for k, pb in res[False][1].items():
    pa = res[True][1][k]
    bad = (pa - pb).abs() > 2e-4 * pb.abs().max().clamp_min(1e-12)
    frac = float(bad.float().mean())
    if frac > worst_frac and bad.numel() > 64:
        worst_frac, worst_key = frac, k
rep["params_fraction_off_worst"] = [worst_key, worst_frac]
rep["loss_rel_diff_max"] = max(abs(a - b) / max(abs(b), 1e-12) for a, b in zip(res[True][0], res[False][0]))
rep["ok"] = bool(res[True][2] and res[False][2] and rep["loss_rel_diff_max"] < 1e-3 and worst_frac <= 0.03)
if rank == 0:
    print(json.dumps(rep))
dist.destroy_process_group()
sys.exit(0 if rep["ok"] else 1)
