"""csrc/dp.cu on ONE GPU: the reduce + Adam + broadcast kernel with a single-rank process group (the rank barriers, the
symmetric-memory plumbing, both the multicast and the peer-pointer form) must reproduce tnf_adam_step bit for bit, and the
ray-count exchange must return what was published.  The multi-rank behaviour (G ranks == 1 rank on the union batch) is
checked on 2 and 8 GPUs by scripts/dp_parity.py (outputs under profiles/r02_dp_parity_*.json)."""
import os
import tempfile

import pytest
import torch
import torch.distributed as dist

pytestmark = pytest.mark.gpu
DEV = "cuda"


@pytest.fixture(scope="module")
def single_rank_group():
    if dist.is_initialized():
        yield
        return
    f = tempfile.NamedTemporaryFile(delete=False)
    f.close()
    torch.cuda.set_device(0)
    dist.init_process_group("nccl", init_method=f"file://{f.name}", rank=0, world_size=1, device_id=torch.device("cuda", 0))
    yield
    dist.destroy_process_group()
    os.unlink(f.name)


@pytest.mark.parametrize("multicast", [True, False])
def test_reduce_adam_bcast_equals_adam_step(single_rank_group, multicast):
    from tinynerf_b200.dp import PeerMemory
    from tinynerf_b200.optim import FusedAdam
    n = 1_000_003
    pm = PeerMemory(n, torch.device("cuda", 0), 0, 1, use_multicast=multicast)
    if multicast and not pm.multicast:
        pytest.skip("no multicast address for a single-rank group on this fabric")
    g = torch.Generator(device="cuda").manual_seed(3)
    p0 = torch.randn(pm.n, device=DEV, generator=g)
    ref = torch.nn.Parameter(p0.clone())
    opt = FusedAdam([ref], lr=1e-2, eps=1e-15, weight_decay=1e-5)
    pm.param.copy_(p0)
    for step in range(1, 4):
        grad = torch.randn(pm.n, device=DEV, generator=g) * 1024.0
        pm.grad.copy_(grad)
        ref.grad = grad.clone()
        # two ranges with their own flag pads, like the planes / heads launches of an iteration
        pm.reduce_adam_bcast(0, 600_000, 0, step, 1e-2, (0.9, 0.999), 1e-15, 1e-5)
        pm.reduce_adam_bcast(600_000, pm.n, 1, step, 1e-2, (0.9, 0.999), 1e-15, 1e-5, n_ctas=7)
        opt.step()
        torch.cuda.synchronize()
        assert int(pm.error.item()) == 0
        assert torch.equal(pm.param, ref.detach()), step
        assert torch.equal(pm.exp_avg, opt.state[ref]["exp_avg"]) and torch.equal(pm.exp_avg_sq, opt.state[ref]["exp_avg_sq"])
        assert torch.equal(pm.grad, grad)          # gradients are only read


def test_ray_count_exchange(single_rank_group):
    from tinynerf_b200.dp import PeerMemory
    pm = PeerMemory(1024, torch.device("cuda", 0), 0, 1)
    for step, cnt in ((1, 17408), (2, 3), (3, 9_000_000), (7, 1)):
        pm.publish_count(step, cnt)
        out = pm.sum_counts(step)
        torch.cuda.synchronize()
        assert float(out) == float(cnt) and int(pm.error.item()) == 0
    pm.check(); pm.check()


def test_argument_validation(single_rank_group):
    from tinynerf_b200.dp import PeerMemory
    pm = PeerMemory(1024, torch.device("cuda", 0), 0, 1)
    with pytest.raises(RuntimeError, match="4-element aligned"):
        from tinynerf_b200 import _lib
        _lib.call("tnf_dp_reduce_adam_bcast", pm._peer_grad, pm._peer_param, None, None, pm.exp_avg.data_ptr(), pm.exp_avg_sq.data_ptr(),
                  2, 10, pm._peer_flags[0], 4, 0, 1, 1, pm.error.data_ptr(), 1e-2, 0.9, 0.999, 1e-15, 0.0, 1, _lib.stream_ptr())
