"""Diagnostics (torchrun): time of the gradient all-reduce alone (132 MB fp32, the K-Planes flat gradient buffer)."""
import os, sys, time
import torch
import torch.distributed as dist

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
for mb in (132, 32, 8):
    x = torch.ones(mb * 1000 * 1000 // 4, device=dev)
    for _ in range(5):
        dist.all_reduce(x)
    torch.cuda.synchronize()
    dist.barrier()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    for _ in range(20):
        dist.all_reduce(x)
    e.record()
    torch.cuda.synchronize()
    ms = s.elapsed_time(e) / 20
    if rank == 0:
        print(f"{os.environ.get('TAG', 'default'):24s} all_reduce {mb:4d} MB x{world}: {ms * 1e3:8.1f} us  algbw {mb / ms:7.1f} GB/s  busbw {mb / ms * 2 * (world - 1) / world:7.1f} GB/s", flush=True)
dist.destroy_process_group()
