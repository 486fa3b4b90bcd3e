"""a15-a17 parity on the GPU: the tcgen05 3xTF32 dense layers (fwd, dgrad, wgrad, fused heads) against fp64
matmuls and against the reference's golden head outputs.  Tolerance 1e-5 relative (north_star) measured
against the natural scale of each dot product (sum_k |x_k w_k|), which is what fp32 SGEMM itself guarantees."""
import pytest
import torch

from tinynerf_b200 import _lib, mlp_ops, models

pytestmark = pytest.mark.gpu
DEV = "cuda"


def lin_fwd(x, w, b, relu):
    y, _ = mlp_ops._lin_fwd(mlp_ops._prep(x), w.contiguous(), b, relu)
    return y


def check(got, want64, scale64, rtol=1e-5):
    err = (got.double() - want64).abs()
    tol = rtol * scale64 + 1e-30
    assert bool((err <= tol).all()), f"worst err/scale {(err / scale64.clamp_min(1e-30)).max().item():.3e}"


@pytest.mark.parametrize("m,k,n", [(1000, 96, 64), (300, 147, 64), (257, 64, 64), (4096, 36, 128), (129, 128, 128), (1, 96, 64),
                                   (70000, 96, 64), (70001, 128, 128), (1, 128, 128), (50000, 36, 128), (3000, 30, 128), (5000, 100, 128)])
@pytest.mark.parametrize("relu", [False, True])
def test_linear_forward(m, k, n, relu):
    g = torch.Generator().manual_seed(m + k)
    x = torch.randn(m, k, generator=g).to(DEV)
    w = (torch.randn(n, k, generator=g) / k ** 0.5).to(DEV)
    b = torch.randn(n, generator=g).to(DEV)
    y = lin_fwd(x, w, b, relu)
    want = x.double() @ w.double().t() + b.double()
    scale = x.double().abs() @ w.double().abs().t() + b.double().abs()
    if relu:
        want = want.clamp_min(0)
    check(y, want, scale)


# the large cases give every CTA several tiles: operand rings wrap, both tensor-memory A sets of the 64-wide weight-gradient
# kernel are reused (K = 148 has room for one set only), K <= 32 shortens its load-ahead distance
@pytest.mark.parametrize("m,k,n", [(1000, 96, 64), (300, 147, 64), (513, 64, 64), (2048, 36, 128), (129, 128, 128),
                                   (200000, 64, 64), (70001, 96, 64), (90000, 148, 64), (50000, 20, 64), (60000, 64, 128),
                                   (70001, 128, 128), (127, 128, 128), (50000, 36, 128), (3000, 30, 128), (5000, 100, 128)])
def test_linear_dgrad_and_wgrad(m, k, n):
    g = torch.Generator().manual_seed(m * 3 + k)
    x = torch.randn(m, k, generator=g).relu().to(DEV)   # an activation: some entries are exactly 0
    w = (torch.randn(n, k, generator=g) / k ** 0.5).to(DEV)
    dy = torch.randn(m, n, generator=g).to(DEV)
    xp = mlp_ops._prep(x)
    ld = (k + 3) // 4 * 4
    dx = torch.empty(m, ld, device=DEV)
    with torch.cuda.device(0):
        _lib.call("tnf_linear_bwd_data", dy.data_ptr(), n, w.data_ptr(), dx.data_ptr(), ld, xp.data_ptr(), xp.stride(0), m, n, k,
                  _lib.stream_ptr())
        gw, gb = torch.zeros_like(w), torch.zeros(n, device=DEV)
        _lib.call("tnf_linear_bwd_weight", dy.data_ptr(), n, xp.data_ptr(), xp.stride(0), gw.data_ptr(), gb.data_ptr(), m, n, k,
                  _lib.stream_ptr())
    want_dx = (dy.double() @ w.double()) * (x > 0)
    check(dx[:, :k], want_dx, dy.double().abs() @ w.double().abs())
    check(gw, dy.double().t() @ x.double(), dy.double().abs().t() @ x.double().abs())
    check(gb, dy.double().sum(0), dy.double().abs().sum(0))


@pytest.mark.parametrize("m", [300, 40000])
def test_weight_stationary_128_layers_vs_resident_kernel(m):
    """128 x 128 layers run with the weights stationary in tensor memory (csrc/wstat.cu); the shared-memory-resident
    linear_kernel (variant 3 = 1) is the same 3xTF32 computation: results agree to fp32 rounding of the accumulation order,
    also without bias / mask and with padded leading dimensions."""
    g = torch.Generator().manual_seed(m)
    ldx, ldy = 132, 136
    xbuf = torch.full((m, ldx), float("nan"))
    xbuf[:, :128] = torch.randn(m, 128, generator=g).relu()
    w = (torch.randn(128, 128, generator=g) / 128 ** 0.5)
    b = torch.randn(128, generator=g)
    xbuf, w, b = xbuf.to(DEV), w.to(DEV), b.to(DEV)
    x = xbuf[:, :128]
    lib = _lib.load()
    outs = {}
    for variant in (0, 1):
        lib.tnf_set_variant(3, variant)
        try:
            y = torch.full((m, ldy), float("nan"), device=DEV)
            dx = torch.full((m, ldy), float("nan"), device=DEV)
            dx0 = torch.full((m, ldy), float("nan"), device=DEV)
            with torch.cuda.device(0):
                _lib.call("tnf_linear_fwd", x.data_ptr(), ldx, w.data_ptr(), None, y.data_ptr(), ldy, m, 128, 128, 0, None, None, None, 0, 0,
                          _lib.stream_ptr())
                _lib.call("tnf_linear_bwd_data", x.data_ptr(), ldx, w.data_ptr(), dx.data_ptr(), ldy, x.data_ptr(), ldx, m, 128, 128,
                          _lib.stream_ptr())
                _lib.call("tnf_linear_bwd_data", x.data_ptr(), ldx, w.data_ptr(), dx0.data_ptr(), ldy, None, 0, m, 128, 128,
                          _lib.stream_ptr())
            outs[variant] = (y, dx, dx0)
        finally:
            lib.tnf_set_variant(3, 0)
    x64, w64 = x.double(), w.double()
    scale_f = x64.abs() @ w64.abs().t()
    scale_b = x64.abs() @ w64.abs()
    for variant in (0, 1):
        y, dx, dx0 = outs[variant]
        assert torch.isnan(y[:, 128:]).all() and torch.isnan(dx[:, 128:]).all()   # nothing written outside the 128 columns
        check(y[:, :128], x64 @ w64.t(), scale_f)
        check(dx[:, :128], (x64 @ w64) * (x > 0), scale_b)
        check(dx0[:, :128], x64 @ w64, scale_b)
    assert torch.equal(outs[0][1][:, :128] == 0, outs[1][1][:, :128] == 0)   # identical ReLU masks


@pytest.mark.parametrize("m,ks", [(1000, (64, 64, 64, 96)), (70001, (64, 64, 64, 96)), (129, (128,)), (50000, (64, 30, 128, 96)),
                                  (200000, (64, 64))])
def test_multi_job_wgrad(m, ks):
    """tnf_linear_bwd_weight_multi: several 64-output weight gradients of one batch in one launch (the operand rings and A
    sets keep counting across the job boundaries, D is flushed per job) == one fp64 matmul per job; twice, accumulating."""
    import ctypes as C
    g = torch.Generator().manual_seed(m + len(ks))
    nj = len(ks)
    dys = [torch.randn(m, 64, generator=g).to(DEV) for _ in ks]
    xs = []
    for k in ks:
        ld = (k + 3) // 4 * 4
        buf = torch.full((m, ld), float("nan"))
        buf[:, :k] = torch.randn(m, k, generator=g).relu()
        xs.append(buf.to(DEV))
    gws = [torch.zeros(64, k, device=DEV) for k in ks]
    gbs = [torch.zeros(64, device=DEV) for _ in ks]
    vp, i64, i32 = (C.c_void_p * nj), (C.c_int64 * nj), (C.c_int32 * nj)
    for rep in (1, 2):
        with torch.cuda.device(0):
            _lib.call("tnf_linear_bwd_weight_multi", nj, vp(*[t.data_ptr() for t in dys]), i64(*[64] * nj), vp(*[t.data_ptr() for t in xs]),
                      i64(*[t.stride(0) for t in xs]), i32(*ks), vp(*[t.data_ptr() for t in gws]), vp(*[t.data_ptr() for t in gbs]), m,
                      _lib.stream_ptr())
        for dy, x, k, gw, gb in zip(dys, xs, ks, gws, gbs):
            x64 = x[:, :k].double()
            check(gw, rep * (dy.double().t() @ x64), rep * (dy.double().abs().t() @ x64.abs()))
            check(gb, rep * dy.double().sum(0), rep * dy.double().abs().sum(0))


@pytest.mark.parametrize("m,ka,kb", [(1000, 51, 96), (70001, 51, 96), (300, 20, 33), (40000, 64, 64)])
def test_wgrad_of_a_concatenated_input(m, ka, kb):
    """tnf_linear_bwd_weight_cat: dW += dy^T [xa | xb] without the concatenation ever being written (the colour head's
    first layer, src/models.py:87)."""
    g = torch.Generator().manual_seed(m + ka)
    xa = torch.full((m, (ka + 3) // 4 * 4), float("nan"))
    xa[:, :ka] = torch.randn(m, ka, generator=g)
    xb = torch.full((m, (kb + 3) // 4 * 4), float("nan"))
    xb[:, :kb] = torch.randn(m, kb, generator=g).relu()
    dy = torch.randn(m, 64, generator=g)
    xa, xb, dy = xa.to(DEV), xb.to(DEV), dy.to(DEV)
    x = torch.cat([xa[:, :ka], xb[:, :kb]], 1).double()
    scratch = torch.zeros(int(_lib.load().tnf_wgrad_cat_scratch_bytes(ka, kb)) // 4, device=DEV)
    for scr in (None, scratch, scratch):   # scalar-atomic flush, then the scratch-tile flush twice (it must come back zeroed)
        gw, gb = torch.zeros(64, ka + kb, device=DEV), torch.zeros(64, device=DEV)
        with torch.cuda.device(0):
            _lib.call("tnf_linear_bwd_weight_cat", dy.data_ptr(), 64, xa.data_ptr(), xa.stride(0), ka, xb.data_ptr(), xb.stride(0), kb,
                      gw.data_ptr(), gb.data_ptr(), m, 64, _lib.ptr(scr), _lib.stream_ptr())
        check(gw, dy.double().t() @ x, dy.double().abs().t() @ x.abs())
        check(gb, dy.double().sum(0), dy.double().abs().sum(0))
        assert not scratch.any()


def test_heads_match_reference_golden(golden):
    g = golden("heads")
    torch.manual_seed(41)
    sig = models.VanillaOpacityDecoder(96).to(DEV)
    col = models.VanillaColorDecoder(8, 96, 64, 3).to(DEV)
    f, d = g["feats"].to(DEV), g["dirs"].to(DEV)
    sig.net.require_supported(f)
    assert torch.allclose(sig(f).cpu(), g["sigma"], rtol=1e-5, atol=1e-7)
    assert torch.allclose(col(f, d).cpu(), g["rgb"], rtol=1e-5, atol=1e-7)


def _safe_rows(mod64, x64, margin=1e-5):
    """Rows whose hidden pre-activations all stay `margin` away from 0 in the fp64 reference.  A ReLU whose
    input is within rounding distance of 0 may switch on/off between any two correct fp32 evaluations (cuBLAS
    vs these kernels vs fp64), which changes that row's gradient by O(1/width): such rows are excluded by
    zeroing their upstream gradient in BOTH evaluations."""
    safe = torch.ones(x64.size(0), dtype=torch.bool, device=x64.device)
    h = x64
    lins = [mm for mm in mod64.modules() if isinstance(mm, torch.nn.Linear)]
    for lin in lins[:-1]:
        pre = torch.nn.functional.linear(h, lin.weight, lin.bias)
        safe &= (pre.abs() > margin).all(1)
        h = pre.relu()
    return safe


@pytest.mark.parametrize("m", [77, 5000, 40000])
def test_fused_heads_forward_backward_vs_torch(m):
    """Whole decoder stacks, forward and every gradient, against the same modules evaluated in float64."""
    import copy
    torch.manual_seed(1)
    sig = models.VanillaOpacityDecoder(96).to(DEV)
    col = models.VanillaColorDecoder(8, 96, 64, 3).to(DEV)
    trunk = models.MLP(36, 128, 5).to(DEV)
    gen = torch.Generator().manual_seed(m)
    f0 = (torch.randn(m, 96, generator=gen) * 0.5).to(DEV)
    d = torch.nn.functional.normalize(torch.randn(m, 3, generator=gen), dim=-1).to(DEV)
    z0 = torch.randn(m, 36, generator=gen).to(DEV)
    sig64, col64, trunk64 = [copy.deepcopy(mm).double() for mm in (sig, col, trunk)]
    f64 = f0.double().requires_grad_(True)
    z64 = z0.double().requires_grad_(True)
    xcol64 = torch.cat([col64.pe(d.double()), d.double(), f64], -1)
    o64 = [torch.exp(sig64.net.net(f64) - 1.0), torch.sigmoid(col64.net.net(xcol64)), trunk64.net(z64)]
    safes = [_safe_rows(sig64.net, f64.detach()), _safe_rows(col64.net, xcol64.detach()), _safe_rows(trunk64, z64.detach())]
    assert all(s.float().mean() > 0.9 for s in safes), [s.float().mean().item() for s in safes]
    f = f0.clone().requires_grad_(True)
    z = z0.clone().requires_grad_(True)
    outs = [sig(f), col(f, d), trunk(z)]
    for a, b in zip(outs, o64):
        assert torch.allclose(a.double(), b, rtol=1e-5, atol=1e-6), (a.double() - b).abs().max()
    gos = [torch.randn_like(o) * s[:, None] for o, s in zip(outs, safes)]
    sum((o * go).sum() for o, go in zip(outs, gos)).backward()
    sum((o * go.double()).sum() for o, go in zip(o64, gos)).backward()
    mine, ref = {"f": f.grad, "z": z.grad}, {"f": f64.grad, "z": z64.grad}
    for name, mod, mod_ref in (("sig", sig, sig64), ("col", col, col64), ("trunk", trunk, trunk64)):
        for (k, p), (_, q) in zip(mod.named_parameters(), mod_ref.named_parameters()):
            mine[f"{name}.{k}"], ref[f"{name}.{k}"] = p.grad, q.grad
    bad = {}
    for k in mine:
        scale = ref[k].abs().max().clamp_min(1e-12)
        e = ((mine[k].double() - ref[k]).abs().max() / scale).item()
        if e > 2e-5:
            bad[k] = e
    assert not bad, bad


def test_color_input_matches_positional_encoding_bitwise(golden):
    """tnf_color_input == cat([PE(d), d, f]) of the reference (src/models.py:33-39,87), bit for bit, plus the
    reference's own PE golden values."""
    g = golden("heads")
    pe = models.PositionalEncoding(8).to(DEV)
    gen = torch.Generator().manual_seed(0)
    d = torch.nn.functional.normalize(torch.randn(5000, 3, generator=gen), dim=-1).to(DEV)
    f = torch.randn(5000, 96, generator=gen).to(DEV).requires_grad_(True)
    x = mlp_ops.color_input(f, d, 8)
    from oracle import ref_port as rp
    want = torch.cat([rp.positional_encoding(d, 8), d, f], -1)   # the reference's formulation with torch's CUDA sin/cos
    assert torch.equal(pe(d), want[:, :48])
    assert x.shape == (5000, 147) and torch.equal(x, want)
    x.backward(torch.ones_like(x))
    assert torch.equal(f.grad, torch.ones_like(f))
    xg = mlp_ops.color_input(g["feats"].to(DEV), g["dirs"].to(DEV), 8)
    assert torch.allclose(xg[:, :48].cpu(), g["pe"], rtol=0, atol=2e-6)  # CPU sinf/cosf of the golden run vs GPU
    # the row without its feature columns ([PE(d) | d | 0], what the fused heads read when the features come from their own
    # rows): packed samples are [N,7] with the directions at columns 3..5
    packed = torch.zeros(5000, 7, device=DEV)
    packed[:, 3:6] = d
    out = torch.full((5000, 52), float("nan"), device=DEV)
    with torch.cuda.device(0):
        _lib.call("tnf_color_input", packed.data_ptr() + 12, 7, None, 0, 8, 0, out.data_ptr(), 52, 5000, _lib.stream_ptr())
    assert torch.equal(out[:, :51], want[:, :51]) and bool((out[:, 51] == 0).all())


@pytest.mark.parametrize("m", [77, 3000])
def test_wide_stacks_forward_backward_vs_float64(m):
    """a17 + the Cobafa colour head: stacks wider than the resident-weight kernels (in > 160 or out > 128) run on the
    streamed-operand tcgen05 kernels (csrc/wide.cu) -- forward and every gradient against float64, same bar as the narrow
    stacks (1e-5 relative on outputs, 3e-5 of the tensor's max on gradients, ReLU-kink rows excluded in both)."""
    import copy
    torch.manual_seed(2)
    trunk = models.VanillaFeatureMLP(10, 256, 8).to(DEV)          # src/run.py:131 : 60 -> 256 (x9) -> 256
    sig = models.VanillaOpacityDecoder(256).to(DEV)               # 256 -> 64 -> 1
    col = models.VanillaColorDecoder(8, 256, 64, 3).to(DEV)       # 307 -> 64 (x4) -> 3
    ccol = models.VanillaColorDecoder(8, 128, 64, 3).to(DEV)      # Cobafa's colour head: 179 -> 64 (x4) -> 3
    gen = torch.Generator().manual_seed(m)
    x0 = (torch.rand(m, 3, generator=gen) * 2 - 1).to(DEV)
    f0 = (torch.randn(m, 256, generator=gen) * 0.3).to(DEV)
    c0 = (torch.randn(m, 128, generator=gen) * 0.3).to(DEV)
    d = torch.nn.functional.normalize(torch.randn(m, 3, generator=gen), dim=-1).to(DEV)
    t64, s64, k64, cc64 = [copy.deepcopy(mm).double() for mm in (trunk, sig, col, ccol)]
    pe10 = trunk.encoding(x0).double()   # fp32 sin/cos of the kernel, shared by both evaluations
    f64 = f0.double().requires_grad_(True)
    c64 = c0.double().requires_grad_(True)
    pe8 = col.pe(d).double()
    xk64 = torch.cat([pe8, d.double(), f64], -1)
    xc64 = torch.cat([pe8, d.double(), c64], -1)
    o64 = [t64.net.net(pe10), torch.exp(s64.net.net(f64) - 1.0), torch.sigmoid(k64.net.net(xk64)), torch.sigmoid(cc64.net.net(xc64))]
    safes = [_safe_rows(t64.net, pe10), _safe_rows(s64.net, f64.detach()), _safe_rows(k64.net, xk64.detach()),
             _safe_rows(cc64.net, xc64.detach())]
    f = f0.clone().requires_grad_(True)
    c = c0.clone().requires_grad_(True)
    outs = [trunk(x0), sig(f), col(f, d), ccol(c, d)]
    for a, b in zip(outs, o64):
        assert a.shape == b.shape
        assert torch.allclose(a.double(), b, rtol=1e-5, atol=1e-6), (a.double() - b).abs().max()
    gos = [torch.randn_like(o) * s[:, None] for o, s in zip(outs, safes)]
    sum((o * go).sum() for o, go in zip(outs, gos)).backward()
    sum((o * go.double()).sum() for o, go in zip(o64, gos)).backward()
    mine, ref = {"f": f.grad, "c": c.grad}, {"f": f64.grad, "c": c64.grad}
    for name, mod, mod_ref in (("trunk", trunk, t64), ("sig", sig, s64), ("col", col, k64), ("ccol", ccol, cc64)):
        for (k, p), (_, q) in zip(mod.named_parameters(), mod_ref.named_parameters()):
            mine[f"{name}.{k}"], ref[f"{name}.{k}"] = p.grad, q.grad
    bad = {}
    for k in mine:
        scale = ref[k].abs().max().clamp_min(1e-12)
        e = ((mine[k].double() - ref[k]).abs().max() / scale).item()
        # measured 1.96e-5 (first trunk layer: ten 256-wide 3xTF32 layers behind it) +- the run-to-run noise of the atomically
        # accumulated weight gradients (2.004e-5 observed): 3e-5 here, inside the documented 5e-5 worst-entry bar
        if e > 3e-5:
            bad[k] = e
    assert not bad, bad
