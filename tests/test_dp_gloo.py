"""Multi-rank host logic on CPU (gloo, world_size 2): ray sharding, union-batch loss normalisation and
gradient all-reduce reproduce the single-process result.  The per-rank "model" is the oracle's CPU
restatement of the renderer (tests may use the oracle); the product's DP plumbing is what is under test."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from tinynerf_b200 import synthetic
from tinynerf_b200.run import RayStore, allreduce_gradients, dp_mse, global_ray_count, shard_slices


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _model(seed):
    torch.manual_seed(seed)
    lin = lambda i, o: [torch.nn.Parameter(torch.randn(o, i) * 0.3), torch.nn.Parameter(torch.zeros(o))]
    return {"feat": [lin(3, 16), lin(16, 16)], "sig": [lin(16, 8), lin(8, 1)], "col": [lin(16 + 27, 8), lin(8, 3)]}


def _render(m, rays_o, rays_d, grid):
    from oracle import ref_port as rp
    aabb = torch.tensor([[-1.5, -1.5, -1.5], [1.5, 1.5, 1.5]])
    packed, info, _ = rp.ray_provider(rays_o, rays_d, grid, 0.01, scene="aabb", n_samples=32, aabb=aabb, near=0.1, far=1e5)
    out = rp.render(lambda x: rp.mlp(m["feat"], x), lambda f: rp.sigma_head(m["sig"], f),
                    lambda f, d: rp.rgb_head(m["col"], 4, f, d), packed, info, torch.ones(3))
    return out


def _params(m):
    return [t for group in m.values() for layer in group for t in layer]


def _worker(rank, world, port, n_rays, out_path):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    torch.set_num_threads(1)
    o, d = synthetic.blender_rays(n_rays, seed=1)
    rgb = torch.rand(n_rays, 3, generator=torch.Generator().manual_seed(2))
    grid = synthetic.analytic_grid(16, seed=3)
    store = RayStore(o, d, rgb, "cpu", seed=7, rank=rank, world=world)
    # ranks take different numbers of rays (dynamic batches): rank 0 -> 40, rank 1 -> 24
    take = 40 if rank == 0 else 24
    ro, rd, target = store.next(take)
    idx = store.last_indices.clone()
    m = _model(5)
    out = _render(m, ro, rd, grid)
    loss = dp_mse(out, target, global_ray_count(take, "cpu", world))
    loss.backward()
    allreduce_gradients(_params(m), world)
    total = loss.detach().clone()
    dist.all_reduce(total)
    gathered = [None] * world
    dist.all_gather_object(gathered, idx.tolist())
    if rank == 0:
        torch.save({"grads": [p.grad.clone() for p in _params(m)], "loss": total, "idx": gathered}, out_path)
    dist.destroy_process_group()


def test_two_rank_step_equals_single_process_on_union_batch(tmp_path):
    n_rays, world = 256, 2
    out_path = tmp_path / "dp.pt"
    mp.spawn(_worker, args=(world, _free_port(), n_rays, str(out_path)), nprocs=world, join=True)
    res = torch.load(out_path)
    idx0, idx1 = res["idx"]
    assert not set(idx0) & set(idx1), "ranks must draw disjoint rays"
    o, d = synthetic.blender_rays(n_rays, seed=1)
    rgb = torch.rand(n_rays, 3, generator=torch.Generator().manual_seed(2))
    grid = synthetic.analytic_grid(16, seed=3)
    union = torch.tensor(idx0 + idx1)
    m = _model(5)
    out = _render(m, o[union], d[union], grid)
    loss = torch.nn.functional.mse_loss(out, rgb[union])
    loss.backward()
    assert float(res["loss"]) == pytest.approx(float(loss), rel=1e-5)
    for a, p in zip(res["grads"], _params(m)):
        assert torch.allclose(a, p.grad, rtol=1e-4, atol=1e-7)


def test_flat_view_of_channels_last_gradients():
    from tinynerf_b200.run import _flat_dense
    g = torch.arange(2 * 3 * 4 * 5, dtype=torch.float32).reshape(1, 6, 4, 5).contiguous(memory_format=torch.channels_last)
    flat = _flat_dense(g)
    assert flat.is_contiguous() and flat.numel() == g.numel() and flat.data_ptr() == g.data_ptr()
    flat.mul_(2.0)
    assert torch.equal(g, torch.arange(120, dtype=torch.float32).reshape(1, 6, 4, 5) * 2)


def test_ray_store_shards_are_a_partition_of_each_epoch():
    o = torch.arange(30, dtype=torch.float32)[:, None].repeat(1, 3)
    stores = [RayStore(o, o, o, "cpu", seed=3, rank=r, world=3) for r in range(3)]
    seen = []
    for st in stores:
        a, _, _ = st.next(10)
        seen += a[:, 0].int().tolist()
    assert sorted(seen) == list(range(30))
    # the order is shuffled lazily (tnf_shuffle_next): over any n consecutive positions every ray appears once, a rewind
    # replays the same rays, and the next epoch is a different order
    st = RayStore(o, o, o, "cpu", seed=5)
    first = st.next(30)[0][:, 0].int().tolist()
    assert sorted(first) == list(range(30)) and first != list(range(30))
    a = st.next(12)[0][:, 0].int().tolist()
    st.rewind(7)
    b = st.next(7)[0][:, 0].int().tolist()
    assert b == a[5:]
    rest = st.next(18)[0][:, 0].int().tolist()
    assert sorted(a + rest) == list(range(30)) and a + rest != first
    assert RayStore(o, o, o, "cpu", seed=5).next(30)[0][:, 0].int().tolist() == first   # seeded
    assert shard_slices(128, 3, 8) == (48, 64)
    with pytest.raises(ValueError):
        shard_slices(128, 0, 3)
