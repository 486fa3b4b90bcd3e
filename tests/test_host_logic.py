"""Host-side mirror of the reference interface (CPU-only checks): module trees / state_dict keys,
channels-last parameter storage, seeded-init parity, marcher/contraction helpers against goldens."""
import torch

from tinynerf_b200 import core, models, synthetic


def test_state_dict_keys_and_shapes_match_reference_layout():
    f = models.KPlanesFeatureField(32)
    keys = list(f.state_dict().keys())
    assert keys == [f"planes.{s}.{p}.plane" for s in range(3) for p in range(3)]
    assert [tuple(v.shape) for v in f.state_dict().values()] == [(1, 32, r, r) for r in (128, 256, 512) for _ in range(3)]
    assert f.feature_dim == 96
    c = models.CobafaFeatureField(torch.linspace(32.0, 128, 6).int().tolist(), 64, torch.linspace(2.0, 8.0, 6).tolist(),
                                  [8, 8, 8, 4, 4, 4], 128)
    k = list(c.state_dict().keys())
    assert k[:7] == [f"basis_grids.{i}.grid" for i in range(6)] + ["coef_grid.grid"]
    assert k[7] == "mlp.net.0.weight" and k[-1] == "mlp.net.7.bias"
    assert sum(p.numel() for p in c.parameters()) == 21_991_356
    d = models.VanillaColorDecoder(8, 96, 64, 3)
    assert list(d.state_dict().keys()) == ["pe.freqs", "net.net.0.weight", "net.net.0.bias", "net.net.2.0.weight",
                                            "net.net.2.0.bias", "net.net.3.0.weight", "net.net.3.0.bias",
                                            "net.net.4.0.weight", "net.net.4.0.bias", "net.net.5.weight", "net.net.5.bias"]


def test_planes_are_channels_last_with_reference_init(golden):
    g = golden("kplanes")
    torch.manual_seed(21)
    f = models.KPlanesFeatureField(32)
    p = f.planes[1][2].plane
    assert p.shape == (1, 32, 256, 256) and p.stride() == (256 * 256 * 32, 1, 256 * 32, 32)
    assert models._channels_last_storage(p).data_ptr() == p.data_ptr()
    chk = float(sum(q.double().sum() for q in f.parameters()))
    assert abs(chk - g["param_checksum"]) < 1e-6 * abs(g["param_checksum"])   # same values as the reference's seeded init
    # (loss_tv / loss_l1 values against the same golden: tests/test_gpu_helpers.py -- the product has no CPU arithmetic)
    # round trip through a contiguous state_dict keeps the kernels' layout
    sd = {k: v.contiguous() for k, v in f.state_dict().items()}
    f.load_state_dict(sd)
    assert models._ensure_channels_last_(f.planes[0][0].plane) is not None


def test_no_cpu_or_eager_path_in_the_product():
    """The mirror's compute methods need CUDA tensors and the built library: on CPU tensors they raise (the reference's
    CHECK_CUDA message), they never fall back to PyTorch arithmetic."""
    import pytest
    pts = torch.randn(10, 3)
    for fn in (lambda: core.ContractionMip360()(pts), lambda: core.ContractionAABB(torch.tensor([[0.0, 0, 0], [2.0, 2, 2]]))(pts),
               lambda: core.RayMarcherAABB(torch.tensor([[-1.0, -1, -1], [1.0, 1, 1]]), 8, 0.1)(pts, pts),
               lambda: models.KPlanesFeaturePlane(8, (16, 16))(pts[:, :2]), lambda: models.KPlanesFeaturePlane(8, (16, 16)).loss_tv(),
               lambda: models.KPlanesFeaturePlane(8, (16, 16)).loss_l1(), lambda: models.CobafaGrid(8, 4)(pts),
               lambda: models.PositionalEncoding(4)(pts), lambda: models.MLP(16, 32, 1, 1)(torch.randn(4, 16)),
               lambda: models.VanillaOpacityDecoder(96)(torch.randn(4, 96))):
        with pytest.raises(RuntimeError, match="must be a CUDA tensor"):
            fn()
    with pytest.raises(NotImplementedError):
        core.ContractionMip360(order=2)(pts)
    # and the product's host modules contain no call of the PyTorch ops the reference's bodies were made of
    import ast
    from tinynerf_b200 import fused, mlp_ops, run
    banned = {"grid_sample", "mse_loss", "index_add_", "index_add", "linear", "relu", "sigmoid", "cumprod", "repeat_interleave"}
    for mod in (core, models, fused, mlp_ops, run):
        tree = ast.parse(open(mod.__file__).read())
        calls = {n.func.attr for n in ast.walk(tree) if isinstance(n, ast.Call) and isinstance(n.func, ast.Attribute)}
        assert not (calls & banned), (mod.__name__, calls & banned)


def test_synthetic_packed_rays_are_a_partition():
    sig, info, g = synthetic.packed_rays(1 << 14, seed=3)
    assert int(info[0, 0]) == 0 and int(info[-1, 0] + info[-1, 1]) == 1 << 14
    assert torch.equal(info[1:, 0], (info[:-1, 0] + info[:-1, 1]))
    assert (info[:, 1] == 0).float().mean() > 0.05 and int(info[:, 1].max()) <= 1024


def test_multistep_lr_matches_torch():
    """run.MultiStepLR (no per-step Python bookkeeping) follows torch's MultiStepLR, repeated milestones included."""
    import torch
    from tinynerf_b200.run import MultiStepLR
    for milestones in ([3, 5, 8], [2, 2, 6], [1]):
        pa, pb = [torch.nn.Parameter(torch.zeros(1))], [torch.nn.Parameter(torch.zeros(1))]
        oa, ob = torch.optim.SGD(pa, lr=1e-2), torch.optim.SGD(pb, lr=1e-2)
        sa, sb = MultiStepLR(oa, milestones, 0.33), torch.optim.lr_scheduler.MultiStepLR(ob, milestones, 0.33)
        for _ in range(12):
            oa.step(); ob.step(); sa.step(); sb.step()
            assert abs(sa.get_last_lr()[0] - sb.get_last_lr()[0]) <= 1e-12 * sb.get_last_lr()[0]


def test_f32_rounding_helper():
    import torch
    from tinynerf_b200.core import _f32
    for v in (0.01, 0.01 ** (1 / 16), 5.196152422706632 / 256, 1e5, 0.1, 1e-45, 3.4e38):
        assert _f32(v) == torch.tensor(v, dtype=torch.float32).item()


def test_partition_and_steps_tags_are_voided_by_in_place_edits():
    """the trusted-partition / contiguous-steps tags must not survive an in-place edit of the tensors (the
    reference's own loop does `info[:, 0] += current_size`, src/run.py:236)."""
    info = torch.tensor([[0, 3], [3, 2]], dtype=torch.int32)
    assert not core.is_trusted_partition(info)
    core.tag_partition(info)
    assert core.is_trusted_partition(info)
    info[:, 0] += 5
    assert not core.is_trusted_partition(info)
    packed, steps = torch.zeros(5, 7), torch.zeros(5)
    assert core.tagged_steps(packed) is None
    core.tag_steps(packed, steps)
    assert core.tagged_steps(packed) is steps
    packed[:, 6] = 1.0
    assert core.tagged_steps(packed) is None
    core.tag_steps(packed, steps)
    steps.add_(1.0)
    assert core.tagged_steps(packed) is None


def test_ray_store_epoch_coverage_with_rewinds():
    """Every ray is handed out exactly once per epoch even when speculative draws are handed back (RayStore.rewind) --
    as long as the caller never rewinds across the epoch boundary, which `remaining()` lets it avoid (Trainer caps the
    chunks it marches at once by it)."""
    from tinynerf_b200.run import RayStore
    n, B = 1000, 64
    o = torch.arange(n, dtype=torch.float32)[:, None].expand(n, 3).contiguous()
    store = RayStore(o, o, o, "cpu", seed=3)
    g = torch.Generator().manual_seed(0)
    for epoch in range(3):
        seen = []
        while True:
            rem = store.remaining()
            K = int(torch.randint(1, 5, (1,), generator=g))
            K = min(K, rem // B) if rem >= B else 1
            draw = min(K * B, rem) if rem < B else K * B
            ro, _, _ = store.next(draw)
            ids = ro[:, 0].long().tolist()
            used = int(torch.randint(1, K + 1, (1,), generator=g)) if K > 1 else K
            if used < K:
                store.rewind((K - used) * B)
                again, _, _ = store.next((K - used) * B)      # replay: the same rays in the same order
                assert again[:, 0].long().tolist() == ids[used * B:]
                store.rewind((K - used) * B)
                ids = ids[:used * B]
            seen += ids
            if store.remaining() == store._m:   # wrapped exactly
                break
        assert sorted(seen) == list(range(n)), f"epoch {epoch}: coverage broken"


def test_trainer_close_restores_gc_state():
    """manual_gc freezes/disables the cyclic collector process-wide; close() must undo that."""
    import gc
    from tinynerf_b200.run import Trainer
    t = Trainer.__new__(Trainer)
    t._gc_frozen, t._gc_was_enabled = True, True
    gc.freeze(); gc.disable()
    t.close()
    assert gc.isenabled() and gc.get_freeze_count() == 0 and not t._gc_frozen


def test_dp_slices_partition_the_parameter_space():
    """dp.slice_of: the ranks' shares of a flat range are contiguous, 4-element aligned (the kernel moves float4), disjoint
    and cover the range -- for ranges that do not divide evenly too."""
    from tinynerf_b200.dp import slice_of
    for lo, hi, world in [(0, 33_030_144, 8), (33_030_144, 33_058_592, 8), (0, 100, 3), (8, 12, 8), (0, 4, 2), (16, 16, 4)]:
        cur = lo
        for r in range(world):
            a, b = slice_of(lo, hi, r, world)
            assert a == cur and a <= b <= hi and (a - lo) % 4 == 0
            cur = b
        assert cur == hi


def test_render_chunk_grouping():
    """run.group_render_chunks (the host half of Trainer.render): chunks are contiguous, cover every ray block once, stay
    within the sample capacity unless a single block exceeds it, and empty blocks are carried along."""
    from tinynerf_b200.run import group_render_chunks
    counts = [0, 0, 500, 900, 100, 0, 2500, 10, 0, 0, 990, 10, 0]
    ends, ray_ends, tot = [], [], 0
    for i, c in enumerate(counts):
        tot += c
        ends.append(tot)
        ray_ends.append(min(2048 * (i + 1), 2048 * len(counts) - 7))
    chunks = group_render_chunks(ends, ray_ends, 1000)
    assert chunks[0][0] == 0 and chunks[0][2] == 0 and chunks[-1][1] == ray_ends[-1] and chunks[-1][3] == tot
    for (a0, a1, s0, s1), (b0, b1, t0, t1) in zip(chunks, chunks[1:]):
        assert a1 == b0 and s1 == t0 and a0 < a1
    sizes = [s1 - s0 for _, _, s0, s1 in chunks]
    assert all(sz <= 1000 or sz == 2500 for sz in sizes) and 2500 in sizes
    assert sum(sizes) == tot
    assert group_render_chunks([0], [5], 1000) == [(0, 5, 0, 0)]
    assert group_render_chunks([], [], 1000) == []
