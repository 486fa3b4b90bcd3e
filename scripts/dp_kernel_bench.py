"""The reduce + Adam + broadcast kernel alone (csrc/dp.cu) under torchrun: time per call over the K-Planes parameter space
for several grid sizes, with NVSwitch multicast and with plain P2P loads/stores; plus tnf_kplanes_bwd-style reductions
into ordinary vs symmetric memory (does peer-mapped / multicast-bound memory slow the scatter's red.v4 down?).

    python -m torch.distributed.run --nproc-per-node G --master-addr 127.0.0.1 --master-port 29620 scripts/dp_kernel_bench.py
"""
import json
import os
import sys
from pathlib import Path

import torch
import torch.distributed as dist

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
from tinynerf_b200 import _lib  # noqa: E402
from tinynerf_b200.dp import PeerMemory  # noqa: E402

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
N = 33_058_592
out = {"world": world, "n_params": N, "runs": []}
sms = torch.cuda.get_device_properties(dev).multi_processor_count
for multicast in (True, False):
    pm = PeerMemory(N, dev, rank, world, use_multicast=multicast)
    pm.grad.normal_()
    pm.param.normal_()
    for n_ctas in (sms // 2, sms, 2 * sms, 4 * sms):
        step = 1
        for _ in range(3):
            pm.reduce_adam_bcast(0, pm.n, 0, step, 1e-2, (0.9, 0.999), 1e-15, 1e-5, n_ctas=n_ctas); step += 1
        torch.cuda.synchronize(); dist.barrier()
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record()
        reps = 10
        for _ in range(reps):
            pm.reduce_adam_bcast(0, pm.n, 0, step, 1e-2, (0.9, 0.999), 1e-15, 1e-5, n_ctas=n_ctas); step += 1
        e.record(); torch.cuda.synchronize()
        us = torch.tensor([s.elapsed_time(e) / reps * 1e3], device=dev)
        dist.all_reduce(us, op=dist.ReduceOp.MAX)
        out["runs"].append({"multicast": bool(pm.multicast), "n_ctas": n_ctas, "us_per_call": round(float(us), 1),
                            "error_word": int(pm.error.item())})
    # NCCL baseline on the same buffers: all-reduce of the gradients + the replicated Adam pass it feeds
    g = pm.grad
    for _ in range(3):
        dist.all_reduce(g)
    torch.cuda.synchronize(); dist.barrier()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    for _ in range(10):
        dist.all_reduce(g)
    e.record(); torch.cuda.synchronize()
    out.setdefault("nccl_allreduce_us", round(s.elapsed_time(e) / 10 * 1e3, 1))
    # red.v4 scatter into the symmetric gradient buffer vs an ordinary allocation (one K-Planes batch)
    if multicast:
        import ctypes as C
        from tinynerf_b200 import core, models, synthetic
        torch.manual_seed(0)
        field = models.KPlanesFeatureField(32).to(dev)
        aabb = torch.tensor([[-1.5, -1.5, -1.5], [1.5, 1.5, 1.5]], device=dev)
        marcher = core.RayMarcherAABB(aabb, 256, 0.1)
        og = core.OccupancyGrid(128, marcher.step_size, 0.01, synthetic.DECAY).to(dev)
        og.grid.copy_(synthetic.analytic_grid(128, seed=1236)); og.mean = og.grid.mean().item()
        prov = core.RayProvider(og, core.ContractionAABB(aabb), marcher)
        o, d = synthetic.blender_rays(9500, seed=2)
        packed, info = prov(o.to(dev), d.to(dev), training=True)
        n = packed.size(0)
        stor = [models._channels_last_storage(p) for p in field._plane_params()]
        ptrs = (C.c_void_p * 9)(*[t.data_ptr() for t in stor])
        res = (C.c_int32 * 3)(128, 256, 512)
        go = torch.randn(n, 96, device=dev)
        sizes = [t.numel() for t in stor]
        flush = torch.empty(64 << 20, device=dev)
        for label, flat in (("ordinary", torch.zeros(sum(sizes), device=dev)), ("symmetric", pm.grad)):
            offs, gp = 0, []
            for sz in sizes:
                gp.append(flat.data_ptr() + 4 * offs); offs += sz
            gptrs = (C.c_void_p * 9)(*gp)
            ts = []
            for _ in range(6):
                flush.zero_()
                s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                s.record()
                _lib.call("tnf_kplanes_bwd", ptrs, gptrs, res, 3, 32, packed.data_ptr(), 7, n, go.data_ptr(), _lib.stream_ptr())
                e.record(); torch.cuda.synchronize()
                ts.append(s.elapsed_time(e) * 1e3)
            out[f"kplanes_bwd_us_{label}_grads"] = round(sorted(ts)[len(ts) // 2], 1)
        # and the gather from parameters in symmetric memory
        pstor = []
        offs = 0
        for t in stor:
            v = pm.param[offs:offs + t.numel()]; v.copy_(t.reshape(-1) if t.is_contiguous() else t.flatten()); pstor.append(v); offs += t.numel()
        pptrs = (C.c_void_p * 9)(*[v.data_ptr() for v in pstor])
        fo = torch.empty(n, 96, device=dev)
        for label, pp in (("ordinary", ptrs), ("symmetric", pptrs)):
            ts = []
            for _ in range(6):
                flush.zero_()
                s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                s.record()
                _lib.call("tnf_kplanes_fwd", pp, res, 3, 32, packed.data_ptr(), 7, n, fo.data_ptr(), _lib.stream_ptr())
                e.record(); torch.cuda.synchronize()
                ts.append(s.elapsed_time(e) * 1e3)
            out[f"kplanes_fwd_us_{label}_planes"] = round(sorted(ts)[len(ts) // 2], 1)
    del pm
    torch.cuda.synchronize(); dist.barrier()
if rank == 0:
    print(json.dumps(out))
dist.destroy_process_group()
