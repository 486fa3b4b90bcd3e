"""tnf_heads_fwd (both decoder heads fused in one tcgen05 kernel) against the per-layer modules (which are themselves
checked against torch fp32 in test_gpu_mlp.py) and against a float64 evaluation.  Tolerance 1e-5 relative (north_star)."""
import ctypes as C

import pytest
import torch

from tinynerf_b200 import _lib, mlp_ops, models

pytestmark = pytest.mark.gpu
DEV = "cuda"


def _run_fused(sig, col, feats, dirs, split=False):
    """split: xc handed to the kernel holds only [PE(d) | d]; the feature columns of the colour input come from `feats`."""
    lib = _lib.load()
    n, F = feats.shape
    xc = mlp_ops.color_input(feats, dirs, col.pe.freqs.numel())
    k0, ld = xc.size(1), xc.stride(0)
    xc_cols = k0 - F if split else k0
    xc_in = xc
    if split:
        xc_in = torch.full((n, (xc_cols + 3) // 4 * 4), float("nan"), device=DEV)   # padding must never be read as data
        xc_in[:, :xc_cols] = xc[:, :xc_cols]
        ld = xc_in.stride(0)
    sl, cl = sig.net.linears(), col.net.linears()
    tab = lambda ts: (C.c_void_p * len(ts))(*[t.data_ptr() for t in ts])
    hs = torch.full((n, 64), float("nan"), device=DEV)
    h = [torch.full((n, 64), float("nan"), device=DEV) for _ in range(4)]
    rgb = torch.full((n, 3), float("nan"), device=DEV)
    sigma = torch.full((n,), float("nan"), device=DEV)
    ws = torch.empty(int(lib.tnf_heads_workspace_bytes(F, k0)) // 4, device=DEV)
    _lib.call("tnf_heads_fwd", feats.data_ptr(), feats.stride(0), F, xc_in.data_ptr(), ld, k0, xc_cols, tab([l.weight for l in cl]),
              tab([l.bias for l in cl]), tab([l.weight for l in sl]), tab([l.bias for l in sl]), tab(h), hs.data_ptr(),
              rgb.data_ptr(), sigma.data_ptr(), n, ws.data_ptr(), _lib.stream_ptr())
    torch.cuda.synchronize()
    return sigma, rgb, hs, h, xc


def _ref64(sig, col, feats, xc):
    f64 = lambda t: t.detach().double()
    sl, cl = sig.net.linears(), col.net.linears()
    hs = torch.relu(f64(feats) @ f64(sl[0].weight).T + f64(sl[0].bias))
    sigma = torch.exp(hs @ f64(sl[1].weight).T + f64(sl[1].bias) - 1.0).ravel()
    x = f64(xc)
    hl = []
    for l in cl[:-1]:
        x = torch.relu(x @ f64(l.weight).T + f64(l.bias))
        hl.append(x)
    rgb = torch.sigmoid(x @ f64(cl[-1].weight).T + f64(cl[-1].bias))
    return sigma, rgb, hs, hl


@pytest.mark.parametrize("split", [False, True])
@pytest.mark.parametrize("n", [1, 127, 128, 129, 300, 148 * 128 + 5, 2 * 148 * 128 + 77, 1 << 18])
def test_heads_fwd_matches_float64_and_per_layer_kernels(n, split):
    torch.manual_seed(n)
    sig = models.VanillaOpacityDecoder(96).to(DEV)
    col = models.VanillaColorDecoder(8, 96, 64, 3).to(DEV)
    feats = torch.randn(n, 96, device=DEV) * 0.5
    dirs = torch.nn.functional.normalize(torch.randn(n, 3, device=DEV), dim=-1)
    sigma, rgb, hs, h, xc = _run_fused(sig, col, feats, dirs, split)
    s64, r64, hs64, h64 = _ref64(sig, col, feats, xc)

    def close(a, b, what):
        err = (a.double() - b).abs()
        tol = 1e-5 * b.abs() + 2e-6
        assert bool((err <= tol).all()), f"{what}: worst excess {(err - tol).max().item():.3e} at n={n}"

    close(hs, hs64, "sigma hidden")
    close(sigma, s64, "sigma")
    for i in range(4):
        close(h[i], h64[i], f"colour hidden {i}")
    close(rgb, r64, "rgb")
    # and the per-layer module path gives the same numbers to fp32 rounding
    with torch.no_grad():
        assert torch.allclose(sig(feats).ravel(), sigma, rtol=2e-5, atol=1e-6)
        assert torch.allclose(col(feats, dirs), rgb, rtol=2e-5, atol=1e-6)


@pytest.mark.parametrize("n", [1, 129, 300, 148 * 128 + 5, 2 * 148 * 128 + 77, 1 << 18])
def test_heads_bwd_data_matches_float64(n):
    """tnf_heads_bwd_data (the data-gradient chain of both heads in one kernel) against a float64 evaluation."""
    torch.manual_seed(1000 + n)
    lib = _lib.load()
    F, K0, col0 = 96, 147, 51
    dh3 = torch.randn(n, 64, device=DEV)
    dhs = torch.randn(n, 64, device=DEV)
    hmask = [torch.randn(n, 64, device=DEV).clamp_min(0.0) for _ in range(3)]        # h2, h1, h0 (post-ReLU: ~half zero)
    W = [torch.randn(64, K0, device=DEV) / 8] + [torch.randn(64, 64, device=DEV) / 8 for _ in range(3)]   # W0..W3
    Ws0 = torch.randn(64, F, device=DEV) / 8
    dh_out = [torch.full((n, 64), float("nan"), device=DEV) for _ in range(3)]
    dfeat = torch.full((n, F), float("nan"), device=DEV)
    ws = torch.empty(int(lib.tnf_heads_bwd_workspace_bytes(F)) // 4, device=DEV)
    tab = lambda ts: (C.c_void_p * len(ts))(*[t.data_ptr() for t in ts])
    _lib.call("tnf_heads_bwd_data", dh3.data_ptr(), dhs.data_ptr(), tab(hmask), tab(W), K0, col0, Ws0.data_ptr(), F, tab(dh_out),
              dfeat.data_ptr(), F, n, ws.data_ptr(), _lib.stream_ptr())
    torch.cuda.synchronize()
    d = dh3.double()
    want = []
    for layer, m in zip((3, 2, 1), hmask):
        d = (d @ W[layer].double()) * (m > 0)
        want.append(d)
    feat = d @ W[0].double()[:, col0:col0 + F] + dhs.double() @ Ws0.double()

    def close(a, b, what):
        err = (a.double() - b).abs()
        tol = 1e-5 * b.abs() + 3e-6 * b.abs().max()   # entries are sums with cancellation: the floor scales with the tensor
        assert bool((err <= tol).all()), f"{what}: worst excess {(err - tol).max().item():.3e} (max |ref| {b.abs().max().item():.3e})"

    for i in range(3):
        close(dh_out[i], want[i], f"dh{2 - i}")
    close(dfeat, feat, "dfeat")
