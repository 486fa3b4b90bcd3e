"""no silent dispatch.  One training iteration of every method of the reference (K-Planes through the
fused step and through the modules, Cobafa, vanilla) and one rendered chunk are run under torch's CUDA profiler; the test
fails if a cuBLAS / CUTLASS GEMM, a torch grid_sampler or an index_add kernel shows up, i.e. if any dense layer or lookup
left the hand-written kernels of libtinynerf_b200.so."""
import pytest
import torch

from tinynerf_b200 import synthetic
from tinynerf_b200.run import RayStore, TrainConfig, Trainer

pytestmark = pytest.mark.gpu
DEV = "cuda"
FORBIDDEN = ("gemm", "cublas", "cutlass", "gemv", "grid_sampler", "index_add", "indexadd", "wmma", "xmma", "splitk")


def _cuda_kernels(fn):
    from torch.autograd import DeviceType
    from torch.profiler import ProfilerActivity, profile
    with profile(activities=[ProfilerActivity.CPU, ProfilerActivity.CUDA]) as prof:
        fn()
        torch.cuda.synchronize()
    return [e.name for e in prof.events() if e.device_type == DeviceType.CUDA]


@pytest.mark.parametrize("method,fused_step,expect", [("kplanes", True, "heads_bwd_data_kernel"), ("kplanes", False, "heads_bwd_data_kernel"),
                                                      ("cobafa", False, "heads_bwd_data_kernel"), ("vanilla", False, "wide")])
def test_training_iteration_runs_only_our_kernels(method, fused_step, expect):
    o, d = synthetic.blender_rays(1 << 13, seed=3)
    rgb = torch.rand(1 << 13, 3, generator=torch.Generator().manual_seed(4))
    torch.manual_seed(7)
    cfg = TrainConfig(method=method, scene_type="aabb", batch_size=256, n_samples=64, fused_step=fused_step, prefetch=False, seed=7)
    tr = Trainer(cfg, RayStore(o, d, rgb, DEV, seed=1), DEV)
    assert (tr._fused is not None) == (method == "kplanes" and fused_step)
    tr.step()   # includes the occupancy update of iteration 0
    names = _cuda_kernels(lambda: (tr.step(), tr.render(o[:600], d[:600], batch_size=256)))
    assert len(names) > 10, "the profiler recorded no CUDA kernels"
    bad = sorted({n for n in names if any(f in n.lower() for f in FORBIDDEN)})
    assert not bad, f"library / torch kernels on the hot path: {bad}"
    ours = [n for n in names if "tnf::" in n]
    assert len(ours) >= 12 and any(expect in n for n in ours), sorted(set(ours))
    tr.close()
