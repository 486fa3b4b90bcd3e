"""The autograd-free training iteration (tinynerf_b200/fused.py) against the module/autograd path on the same batch,
plus the two small kernels it adds (tnf_mse_loss_grad, tnf_tv_fwd_bwd).  All through the C ABI on the GPU."""
import ctypes as C

import pytest
import torch

from tinynerf_b200 import _lib, models, synthetic

pytestmark = pytest.mark.gpu
DEV = "cuda"


def _trainer(fused_step, seed=9, method="kplanes", **kw):
    from tinynerf_b200.run import RayStore, TrainConfig, Trainer
    o, d = synthetic.blender_rays(1 << 15, seed=3)
    rgb = torch.rand(1 << 15, 3, generator=torch.Generator().manual_seed(4))
    cfg = TrainConfig(method=method, scene_type="aabb", batch_size=256, n_samples=128, fused_tv_grad=False,
                      prefetch=False, fused_step=fused_step, **kw)
    torch.manual_seed(seed)
    tr = Trainer(cfg, RayStore(o, d, rgb, DEV, seed=1), DEV)
    tr.occupancy_grid.grid.copy_(synthetic.analytic_grid(128, seed=5).to(DEV))
    tr.occupancy_grid.mean = tr.occupancy_grid.grid.mean().item()
    tr.train_step = 1  # skip the occupancy update of step 0
    return tr


def test_fused_step_matches_autograd_step():
    res = {}
    for fused in (False, True):
        tr = _trainer(fused)
        assert (tr._fused is not None) == fused
        tr.optimizer.step = lambda *a, **k: None  # keep this step's gradients for inspection
        torch.manual_seed(10)
        info = tr.step()
        res[fused] = ({k: p.grad.clone() for k, p in tr.renderer.named_parameters()}, float(info["loss"]), info["n_samples"])
    assert res[True][2] == res[False][2] > 0
    assert res[True][1] == pytest.approx(res[False][1], rel=2e-6)
    for k, b in res[False][0].items():
        a = res[True][0][k]
        assert a.shape == b.shape and a.stride() == b.stride(), k
        # 1e-5 of the tensor's scale, except where a ReLU flipped: the two paths evaluate the heads with different (equally
        # valid) fp32 summation orders, so among the ~10^7 hidden pre-activations of a batch a few lie within rounding
        # distance of zero and switch on in one path only; such a sample moves the entries it touches by its own share of the
        # gradient (~1/n_samples).  Those entries must stay rare (0.1 %; 10 % for the small head tensors, where a flipped
        # unit moves a whole row or several of 64 bias entries) and small (1e-3 of the scale).
        scale = b.abs().max().clamp_min(1e-12)
        err = (a - b).abs()
        assert err.max() <= 1e-3 * scale, k
        assert float((err > 1e-5 * scale).float().mean()) <= (1e-3 if a.numel() > 100000 else 0.1), k


def test_fused_cobafa_step_matches_autograd_step():
    """fused_cobafa.FusedCobafaStep against the module/autograd path on the same batch and the same generator state (the
    Dropout(0.01) mask is drawn by the same kernel from the same stream in both): loss and every gradient, same bars as the
    K-Planes test above."""
    res = {}
    for fused in (False, True):
        tr = _trainer(fused, method="cobafa")
        assert (tr._fused_cobafa is not None) == fused and tr._fused is None
        tr.optimizer.step = lambda *a, **k: None  # keep this step's gradients for inspection
        torch.manual_seed(10)
        info = tr.step()
        res[fused] = ({k: p.grad.clone() for k, p in tr.renderer.named_parameters()}, float(info["loss"]), info["n_samples"])
        tr.close()
    assert res[True][2] == res[False][2] > 0
    assert res[True][1] == pytest.approx(res[False][1], rel=2e-6)
    for k, b in res[False][0].items():
        a = res[True][0][k]
        assert a.shape == b.shape and a.stride() == b.stride(), k
        scale = b.abs().max().clamp_min(1e-12)
        err = (a - b).abs()
        assert err.max() <= 1e-3 * scale, (k, float(err.max() / scale))
        assert float((err > 1e-5 * scale).float().mean()) <= (1e-3 if a.numel() > 100000 else 0.1), k


def test_fused_cobafa_density_matches_modules():
    """FusedCobafaStep.density (the occupancy update's sigma_fn) == sigma_decoder(feature_module(x)) of the modules, in
    training mode (same dropout mask from the same generator state) and in eval mode."""
    tr = _trainer(True, method="cobafa")
    x = (torch.rand(20000, 3, device=DEV) * 2 - 1).contiguous()
    for training in (True, False):
        tr.renderer.train(training)
        torch.manual_seed(77)
        with torch.no_grad():
            want = tr.renderer.sigma_decoder(tr.renderer.feature_module(x))
        torch.manual_seed(77)
        got = tr._fused_cobafa.density(x)
        assert got.shape == want.shape
        assert torch.allclose(got, want, rtol=1e-5, atol=1e-7), float((got - want).abs().max())
    tr.close()


def test_fused_cobafa_training_runs_and_learns():
    """A few full iterations (Adam, scheduler, an occupancy update through the modules in between): finite, decreasing loss,
    and the fused step survives optimizer.zero_grad() and a workspace re-allocation."""
    tr = _trainer(True, seed=12, method="cobafa")
    tr.train_step = 0   # include the occupancy update of iteration 0 (module path's sigma_fn)
    losses = []
    for it in range(6):
        torch.manual_seed(200 + it)
        losses.append(float(tr.step()["loss"]))
        if it == 2:
            tr.optimizer.zero_grad()
            tr._fused_cobafa._cap_n = tr._fused_cobafa._cap_r = 0
    assert all(l == l and l < 1e3 for l in losses)
    assert losses[-1] < losses[0]
    tr.close()


def test_fused_training_trajectory_matches_autograd():
    """Five full iterations (Adam included): parameters stay together.  Adam's update is ~lr*sign(g) where |g| >> eps, so
    an entry whose gradient is a rounding-level residue (float atomics order differs run to run) may step the other
    way in either path -- and a hidden unit that is almost dead (its pre-activation crosses zero for a handful of samples,
    decided differently by the two forward kernels' rounding) does so for its whole weight row.  The test therefore bounds
    the FRACTION of entries that drift (3 %: two rows of the widest layer), not the single worst one."""
    trs = {f: _trainer(f, seed=11) for f in (False, True)}
    for it in range(5):
        for f, tr in trs.items():
            torch.manual_seed(100 + it)
            tr.step()
    pa, pb = dict(trs[True].renderer.named_parameters()), dict(trs[False].renderer.named_parameters())
    for k in pb:
        bad = (pa[k] - pb[k]).abs() > 2e-4 * pb[k].abs().max().clamp_min(1e-12)
        allowed = max(0.03, 2.0 / bad.numel())  # small tensors: up to two stray entries
        assert float(bad.float().mean()) <= allowed, (k, float(bad.float().mean()))
    assert float(trs[True].last["loss"]) == pytest.approx(float(trs[False].last["loss"]), rel=1e-3)


def test_fused_step_survives_zero_grad_and_growth():
    tr = _trainer(True)
    tr.step()
    tr.optimizer.zero_grad()  # set_to_none=True: the fused step re-attaches its gradient views
    tr._fused._cap_n = tr._fused._cap_r = 0  # force a workspace re-allocation
    out = tr.step()
    assert torch.isfinite(out["loss"]) and all(p.grad is not None for p in tr.renderer.parameters())


def test_loss_ring_returns_every_iterations_loss():
    tr = _trainer(True)
    seen = []
    for it in range(5):
        seen.append(float(tr.step()["loss"]))
        if it >= 1:  # one iteration late, as a logging loop would
            assert tr.read_loss(tr.train_step - 2) == seen[-2]
    assert tr.read_loss() == seen[-1]
    with pytest.raises(ValueError):
        tr.read_loss(tr.train_step)


def test_mse_loss_grad_kernel():
    torch.manual_seed(0)
    for r, n_glob in ((1, None), (1000, None), (4097, 9000.0)):
        a = torch.rand(r, 3, device=DEV, requires_grad=True)
        t = torch.rand(r, 3, device=DEV)
        denom = (n_glob or r) * 3
        loss_ref = ((a - t) ** 2).sum() / denom
        (loss_ref * 1024.0).backward()
        g, l = torch.empty(r, 3, device=DEV), torch.zeros(1, device=DEV)
        ng = None if n_glob is None else torch.tensor(n_glob, device=DEV)
        _lib.call("tnf_mse_loss_grad", a.data_ptr(), t.data_ptr(), r, float(r), _lib.ptr(ng), 1024.0, g.data_ptr(), l.data_ptr(),
                  _lib.stream_ptr())
        assert float(l) == pytest.approx(float(loss_ref), rel=2e-6)
        assert torch.allclose(g, a.grad, rtol=2e-6, atol=0)


def test_composite_loss_kernel_equals_the_three_separate_kernels():
    sig, info, _ = synthetic.packed_rays(1 << 16, seed=5, mean_len=40, max_len=300)
    n, r = sig.numel(), info.size(0)
    g = torch.Generator().manual_seed(6)
    w = (torch.rand(n, generator=g) * (torch.rand(n, generator=g) > 0.3)).to(DEV)     # some weights are exactly 0
    rgb, target = torch.rand(n, 3, generator=g).to(DEV), torch.rand(r, 3, generator=g).to(DEV)
    info = info.to(DEV)
    bg = (C.c_float * 3)(1.0, 0.5, 0.25)
    st = _lib.stream_ptr()
    scratch = torch.zeros(2, dtype=torch.float64, device=DEV)
    for n_glob in (None, torch.tensor(2.5 * r, device=DEV)):
        out_a, go, gw_a, grgb_a, loss_a = (torch.empty(r, 3, device=DEV), torch.empty(r, 3, device=DEV), torch.empty(n, device=DEV),
                                           torch.empty(n, 3, device=DEV), torch.zeros(1, device=DEV))
        _lib.call("tnf_composite_fwd", w.data_ptr(), rgb.data_ptr(), info.data_ptr(), n, r, bg, out_a.data_ptr(), None, st)
        _lib.call("tnf_mse_loss_grad", out_a.data_ptr(), target.data_ptr(), r, float(r), _lib.ptr(n_glob), 1024.0, go.data_ptr(),
                  loss_a.data_ptr(), st)
        _lib.call("tnf_composite_bwd", w.data_ptr(), rgb.data_ptr(), info.data_ptr(), n, r, bg, go.data_ptr(), gw_a.data_ptr(),
                  grgb_a.data_ptr(), st)
        out_b, gw_b, grgb_b, loss_b = torch.empty(r, 3, device=DEV), torch.empty(n, device=DEV), torch.empty(n, 3, device=DEV), torch.zeros(1, device=DEV)
        for _ in range(2):   # twice: the scratch words must come back zeroed
            _lib.call("tnf_composite_loss_fwd_bwd", w.data_ptr(), rgb.data_ptr(), info.data_ptr(), n, r, bg, target.data_ptr(), float(r),
                      _lib.ptr(n_glob), 1024.0, out_b.data_ptr(), gw_b.data_ptr(), grgb_b.data_ptr(), loss_b.data_ptr(),
                      scratch.data_ptr(), None, None, 0, st)
            assert torch.equal(out_a, out_b) and torch.equal(gw_a, gw_b) and torch.equal(grgb_a, grgb_b)
            assert float(loss_b) == pytest.approx(float(loss_a), rel=1e-6)
            assert not scratch.any()
        # extra loss terms (the weighted TV sums in the training iteration) are added to the reported value only
        terms = torch.tensor([0.5, 2.0, 3.0], dtype=torch.float64, device=DEV)
        coef = torch.tensor([0.25, 0.125, 1.0], dtype=torch.float64, device=DEV)
        _lib.call("tnf_composite_loss_fwd_bwd", w.data_ptr(), rgb.data_ptr(), info.data_ptr(), n, r, bg, target.data_ptr(), float(r),
                  _lib.ptr(n_glob), 1024.0, out_b.data_ptr(), gw_b.data_ptr(), grgb_b.data_ptr(), loss_b.data_ptr(),
                  scratch.data_ptr(), terms.data_ptr(), coef.data_ptr(), 3, st)
        assert float(loss_b) == pytest.approx(float(loss_a) + 3.375, rel=1e-6) and torch.equal(gw_a, gw_b)


def test_tv_fwd_bwd_equals_separate_passes():
    torch.manual_seed(1)
    field = models.KPlanesFeatureField(32).to(DEV)
    params = field._plane_params()
    n = len(params)
    stor = [models._channels_last_storage(p) for p in params]
    ptrs = (C.c_void_p * n)(*[t.data_ptr() for t in stor])
    res = (C.c_int32 * n)(*[int(p.shape[-1]) for p in params])
    wts = (C.c_float * n)(*[1.0 / n] * n)
    gs = torch.full((1,), 0.37, device=DEV)
    sums_a = torch.empty(2 * n, dtype=torch.float64, device=DEV)
    sums_b = torch.empty(2 * n, dtype=torch.float64, device=DEV)
    base = [torch.randn_like(t) for t in stor]
    ga, gb = [b.clone() for b in base], [b.clone() for b in base]
    st = _lib.stream_ptr()
    _lib.call("tnf_tv_fwd", ptrs, res, n, 32, sums_a.data_ptr(), st)
    _lib.call("tnf_tv_bwd", ptrs, (C.c_void_p * n)(*[t.data_ptr() for t in ga]), res, n, 32, wts, gs.data_ptr(), 1, st)
    _lib.call("tnf_tv_fwd_bwd", ptrs, (C.c_void_p * n)(*[t.data_ptr() for t in gb]), res, n, 32, wts, gs.data_ptr(), 1,
              sums_b.data_ptr(), st)
    assert torch.allclose(sums_a, sums_b, rtol=1e-6, atol=0)   # fp32 partial sums per block vs double per thread
    for x, y in zip(ga, gb):
        assert torch.equal(x, y)
    # the row-marching kernel (default for 32-channel planes) and the one-thread-per-texel kernel agree bit for bit on the
    # gradient, in both the accumulate and the overwrite mode
    import os
    for acc in (1, 0):
        gc_, gd = [b.clone() for b in base], [b.clone() for b in base]
        sums_c = torch.empty_like(sums_b)
        _lib.load().tnf_set_variant(1, 1)   # the one-thread-per-texel kernel
        try:
            _lib.call("tnf_tv_fwd_bwd", ptrs, (C.c_void_p * n)(*[t.data_ptr() for t in gc_]), res, n, 32, wts, gs.data_ptr(), acc,
                      sums_c.data_ptr(), st)
        finally:
            _lib.load().tnf_set_variant(1, 0)
        _lib.call("tnf_tv_fwd_bwd", ptrs, (C.c_void_p * n)(*[t.data_ptr() for t in gd]), res, n, 32, wts, gs.data_ptr(), acc,
                  sums_b.data_ptr(), st)
        assert torch.allclose(sums_c, sums_b, rtol=1e-6, atol=0)
        for x, y in zip(gc_, gd):
            assert torch.equal(x, y)
    # and the value is the reference's loss_tv
    ref = field.loss_tv()
    denom = torch.tensor([float(32 * (r - 1) * r) for r in res for _ in range(2)], dtype=torch.float64, device=DEV)
    assert float((sums_b / denom).sum() / n) == pytest.approx(float(ref), rel=1e-6)


def test_fused_step_vs_reference_restatement():
    """the path bench.py times (FusedKPlanesStep.forward_backward: tnf_heads_fwd, tnf_heads_bwd_data,
    the TMA weight-gradient kernels, split colour input, tnf_composite_loss_fwd_bwd, TV written into the gradients)
    DIRECTLY against the reference restatement (oracle/ref_port.py on the GPU: stock grid_sample / Linear / index_add_ +
    the UNMODIFIED reference weights kernel) on a ~2^18-sample batch -- no autograd path of ours in between.  The
    measurement lives in oracle/parity.py (bench.py prints the same numbers as its `parity` object).
    Bar: rendered colours and loss 1e-5 relative; every parameter gradient rel-L2 <= 2e-5 and worst entry <= 5e-5 of the
    tensor's max (same bar as test_gpu_models.test_kplanes_renderer_vs_torch_on_gpu).  Rays holding a sample whose hidden
    pre-activation lies within 3e-6 of a ReLU kink (decided differently by two correct fp32 evaluations) are removed
    from the batch of BOTH pipelines."""
    from oracle import parity
    r = parity.fused_step_parity()
    print("fused-vs-reference:", {k: v for k, v in r.items() if k != "per_parameter"})
    assert r["n_samples"] > 200_000 and r["kink_rays_excluded_frac"] < 0.25
    assert r["rendered_excess_over_1e-5rel+2e-6"] <= 0.0, r["rendered_max_abs_err"]
    assert r["loss_rel_err"] <= 1e-5
    bad = {k: v for k, v in r["per_parameter"].items() if v[0] > 2e-5 or v[1] > 5e-5}
    assert not bad, bad


def test_parity_object_of_the_bench_line():
    """The march and weights entries of bench.py's `parity` object: bit-exact flags true, errors inside the stated bars."""
    from oracle import parity
    m = parity.march_parity()
    assert m["packing_info_bit_exact"] and m["packed_rows_bit_exact"] and m["n_samples"] > 50_000
    w = parity.weights_parity()
    assert w["termination_mask_bit_exact"] and w["weights_max_rel_err"] <= 1e-5 and w["grad_sigmas_max_err_over_bound"] <= 1.0


def test_sorted_plane_scatter_equals_direct_scatter():
    """tnf_kplanes_sort + tnf_kplanes_bwd_sorted (coarse scales reduced run by run in each plane's own 2-D order) against
    tnf_kplanes_bwd (one reduction per sample and corner) on a ~2^18-sample batch: same plane gradients up to fp32
    summation order; the sort tables are permutations and carry the right coordinates."""
    from oracle import parity
    renderer, prov, og, aabb, o, d = parity.kplanes_case(9600)
    torch.manual_seed(5)
    packed, info = prov(o, d, training=True)
    n = packed.size(0)
    field = renderer.feature_module
    planes = field._plane_params()
    stor = [models._channels_last_storage(p) for p in planes]
    ptrs = (C.c_void_p * 9)(*[t.data_ptr() for t in stor])
    res = (C.c_int32 * 3)(128, 256, 512)
    go = torch.randn(n, 96, device=DEV)
    st = _lib.stream_ptr()
    lib = _lib.load()
    direct = [torch.zeros_like(t) for t in stor]
    _lib.call("tnf_kplanes_bwd", ptrs, (C.c_void_p * 9)(*[t.data_ptr() for t in direct]), res, 3, 32, packed.data_ptr(), 7, n,
              go.data_ptr(), st)
    pos = torch.empty(3, n, dtype=torch.int32, device=DEV)
    uv = torch.empty(3, n, 2, device=DEV)
    scratch = torch.empty(int(lib.tnf_kplanes_sort_scratch_ints(256, n)), dtype=torch.int32, device=DEV)
    _lib.call("tnf_kplanes_sort", packed.data_ptr(), 7, n, 256, scratch.data_ptr(), pos.data_ptr(), uv.data_ptr(), st)
    torch.cuda.synchronize()
    for o_, (a, b) in enumerate(((0, 1), (0, 2), (1, 2))):
        assert torch.equal(pos[o_].long().sort().values, torch.arange(n, device=DEV))          # a permutation
        assert torch.equal(uv[o_][pos[o_].long()], packed[:, [a, b]])                            # slot holds its sample's coordinates
        cell = ((uv[o_] + 1) * 0.5 * 255).floor().long().clamp(0, 255)
        key = (((cell[:, 1] >> 1) * 128 + (cell[:, 0] >> 1)) << 2) | ((cell[:, 1] & 1) << 1) | (cell[:, 0] & 1)
        assert bool((key[1:] >= key[:-1]).all())                                                # slot order is sorted by cell
    rows = torch.empty(6 * n * 32, device=DEV)
    for phases in ((0,), (1, 2)):
        got = [torch.zeros_like(t) for t in stor]
        gp = (C.c_void_p * 9)(*[t.data_ptr() for t in got])
        for ph in phases:
            _lib.call("tnf_kplanes_bwd_sorted", ptrs, gp, res, 3, 32, packed.data_ptr(), 7, n, go.data_ptr(), 2, pos.data_ptr(),
                      uv.data_ptr(), rows.data_ptr(), ph, st)
        for i, (a, b) in enumerate(zip(got, direct)):
            scale = b.abs().max()
            assert float((a - b).abs().max()) <= 2e-5 * float(scale), (phases, i)
            assert float((a - b).norm() / b.norm()) <= 2e-6, (phases, i)
    # and through the trainer (opt-in): a batch drawn by the trainer is tagged, the fused step takes the sorted path
    import os
    os.environ["TNF_KPLANES_SORTED"] = "1"
    try:
        tr = _trainer(True, seed=13)
    finally:
        del os.environ["TNF_KPLANES_SORTED"]
    batch = tr.next_batch()
    assert tr._fused.sorted_scales == 2 and tr._fused.sorted_tag(batch[0]) is not None
    out_a = tr._fused.forward_backward(batch[0], batch[2], batch[1])
    ga = {k: p.grad.clone() for k, p in tr.renderer.named_parameters()}
    del batch[0]._tnf_ksort
    out_b = tr._fused.forward_backward(batch[0], batch[2], batch[1])
    assert float(out_a["loss"]) == float(out_b["loss"])
    for k, p in tr.renderer.named_parameters():
        sc = p.grad.abs().max().clamp_min(1e-12)
        assert float((ga[k] - p.grad).abs().max()) <= 2e-5 * float(sc), k
