"""The oracle (oracle/tnf_oracle.c + oracle/ref_port.py) against the reference's own outputs
(tests/golden/*.npz, produced by importing /root/reference -- tests/golden/make_golden.py) and against
the reference's only known-answer test.  CPU only."""
import math

import pytest
import torch

from oracle import c as orc
from oracle import ref_port as rp
from tinynerf_b200 import synthetic


def test_reference_kat_occupancy_axis_order():
    """tests/test_core.py:5-38 of the reference: pins coord x <-> last grid dim."""
    grid = torch.ones(128, 128, 128)
    grid[:, :, 64:] = 0.0
    pts = [[32, 32, 32], [32, 32, 96], [32, 96, 32], [32, 96, 96], [96, 32, 32], [96, 32, 96], [96, 96, 32], [96, 96, 96]]
    unit = 2.0 * (torch.tensor(pts) / torch.tensor([128.0, 128.0, 128.0])) - 1.0
    want = torch.tensor([True] * 4 + [False] * 4)
    assert torch.equal(orc.occ_query(grid, unit, 0.01)[0], want)
    assert torch.equal(rp.occupancy_lookup(grid, unit, 0.01), want)


@pytest.mark.parametrize("scene", ["aabb", "unbounded"])
@pytest.mark.parametrize("training", [False, True])
def test_provider_matches_reference(golden, scene, training):
    g = golden(f"provider_{scene}")
    noise = g["noise"] if training else None
    want_p, want_i = (g["packed_train"], g["info_train"]) if training else (g["packed_eval"], g["info_eval"])
    kw = dict(scene=scene, n_samples=64, aabb=g["aabb"], near=0.1, far=1e5, uniform_range=1.7, noise=noise)
    packed, info, _ = rp.ray_provider(g["rays_o"], g["rays_d"], g["grid"], g["threshold"], **kw)
    assert torch.equal(info, want_i)
    assert torch.equal(packed, want_p)
    # C restatement: same mask / info / packed rows, bit for bit
    if scene == "aabb":
        m, i, p = orc.march(0, g["rays_o"], g["rays_d"], 64, g["grid"], g["threshold"], aabb=g["aabb"], near=0.1,
                            far=1e5, step_size=g["step_size"], noise=noise)
    else:
        tt, ss = rp.unbounded_tables(64, 0.1, 1.7, "cpu")
        m, i, p = orc.march(1, g["rays_o"], g["rays_d"], 64, g["grid"], g["threshold"], t_table=tt, step_table=ss,
                            noise=noise)
    assert torch.equal(i, want_i)
    assert torch.equal(p, want_p)


def test_marcher_tables_match_reference(golden):
    g = golden("provider_unbounded")
    tt, ss = rp.unbounded_tables(64, 0.1, 1.7, "cpu")
    assert torch.equal(tt.expand(48, 64), g["t_values"]) and torch.equal(ss.expand(48, 64), g["step_sizes"])
    g = golden("provider_aabb")
    step = rp.aabb_step_size(g["aabb"], 64)
    assert float(step) == g["step_size"]
    t = rp.aabb_t_min(g["rays_o"], g["rays_d"], g["aabb"], 0.1, 1e5)[:, None] + torch.arange(64, dtype=torch.float) * step
    assert torch.equal(t, g["t_values"])


def test_occupancy_update_matches_reference(golden):
    g = golden("occ_update")
    sigma_fn = lambda x: 60.0 * torch.exp(-6.0 * (x ** 2).sum(-1, keepdim=True))
    new, mean = rp.occupancy_update(g["grid_before"], g["noise"], sigma_fn, g["step_size"], g["threshold_before"], g["decay"])
    assert torch.equal(new, g["grid_after"]) and mean == g["mean_after"]
    # C restatement of the two elementwise halves around sigma_fn
    coords = orc.occ_update_coords((16, 16, 16), 0, g["noise"])
    sig = sigma_fn(coords)
    new_c = orc.occ_update_apply(g["grid_before"].reshape(-1), 0, sig, g["step_size"],
                                 float(torch.tensor(g["threshold_before"], dtype=torch.float32)),
                                 float(torch.tensor(g["decay"], dtype=torch.float32)))
    assert torch.equal(new_c.view(16, 16, 16), g["grid_after"])


def _planes_from_seed(seed):
    torch.manual_seed(seed)
    return [[torch.nn.init.uniform_(torch.empty(1, 32, r, r)) for _ in range(3)] for r in (128, 256, 512)]


def test_kplanes_matches_reference(golden):
    g = golden("kplanes")
    planes = _planes_from_seed(21)
    for s in planes:
        for p in s:
            p.requires_grad_(True)
    feats = rp.kplanes_features(planes, g["x"])
    assert torch.equal(feats, g["features"])
    (feats * g["grad_out"]).sum().backward()
    for s in range(3):
        for p in range(3):
            gr = planes[s][p].grad.reshape(-1)
            assert torch.equal(gr[g[f"gidx_{s}_{p}"]], g[f"gval_{s}_{p}"])
    assert abs(float(rp.kplanes_tv(planes).detach()) - g["tv"]) <= 1e-6 * abs(g["tv"])


def test_cobafa_matches_reference(golden):
    g = golden("cobafa")
    torch.manual_seed(31)
    basis = [torch.nn.init.uniform_(torch.empty(1, c, r, r, r)) for r, c in zip([8, 11, 14], [8, 4, 2])]
    coef = torch.nn.init.uniform_(torch.empty(1, 3, 9, 9, 9))
    freqs = torch.linspace(2.0, 8.0, 3).tolist()
    for t in basis + [coef]:
        t.requires_grad_(True)
    lookup = rp.cobafa_lookup(basis, coef, freqs, g["x"])
    assert torch.equal(lookup, g["lookup"])
    (lookup * g["grad_out"]).sum().backward()
    for i in range(3):
        assert torch.equal(basis[i].grad, g[f"gbasis_{i}"])
    assert torch.equal(coef.grad, g["gcoef"])


def _linear_params(seed_modules):
    return [(m.weight, m.bias) for m in seed_modules]


def test_heads_match_reference(golden):
    from tinynerf_b200 import models  # host-side classes: same init order => same parameters
    g = golden("heads")
    torch.manual_seed(41)
    sig = models.VanillaOpacityDecoder(96)
    col = models.VanillaColorDecoder(8, 96, 64, 3)
    s_layers = [(l.weight, l.bias) for l in sig.net.linears()]
    c_layers = [(l.weight, l.bias) for l in col.net.linears()]
    assert torch.equal(rp.positional_encoding(g["dirs"], 8), g["pe"])
    assert torch.allclose(rp.sigma_head(s_layers, g["feats"]), g["sigma"], rtol=1e-6, atol=1e-7)
    assert torch.allclose(rp.rgb_head(c_layers, 8, g["feats"], g["dirs"]), g["rgb"], rtol=1e-6, atol=1e-7)
    # (the host-side module classes run on the GPU only -- tests/test_gpu_mlp.py::test_heads_match_reference_golden)


def test_renderer_matches_reference(golden):
    from tinynerf_b200 import models
    g = golden("renderer")
    torch.manual_seed(51)
    fm = models.VanillaFeatureMLP(4, 32, 1)
    sd = models.VanillaOpacityDecoder(32)
    cd = models.VanillaColorDecoder(4, 32, 32, 1)
    with torch.no_grad():
        sd.net.net[-1].bias += 5.0
    aabb = torch.tensor([[-1.5, -1.5, -1.5], [1.5, 1.5, 1.5]])
    thr = min(0.01, g["grid"].mean().item())
    packed, info, _ = rp.ray_provider(g["rays_o"], g["rays_d"], g["grid"], thr, scene="aabb", n_samples=48, aabb=aabb,
                                      near=0.1, far=1e5)
    assert torch.equal(packed, g["packed"]) and torch.equal(info, g["info"])
    f_layers = [(l.weight, l.bias) for l in fm.net.linears()]
    s_layers = [(l.weight, l.bias) for l in sd.net.linears()]
    c_layers = [(l.weight, l.bias) for l in cd.net.linears()]
    out, aux = rp.render(lambda x: rp.mlp(f_layers, rp.positional_encoding(x, 4)), lambda f: rp.sigma_head(s_layers, f),
                         lambda f, d: rp.rgb_head(c_layers, 4, f, d), packed, info, torch.ones(3), return_aux=True)
    assert (aux["weights"] == 0).any(), "fixture should exercise early termination"
    assert torch.allclose(out, g["rendered"], rtol=1e-6, atol=1e-7)
    loss = ((out - 0.25) ** 2).mean()
    loss.backward()
    assert torch.allclose(sd.net.net[0].weight.grad, g["grad_sigma_w"], rtol=1e-4, atol=1e-8)
    assert torch.allclose(fm.net.net[0].weight.grad, g["grad_feat_w"], rtol=1e-4, atol=1e-8)


def _weights_autograd_f64(sigmas, steps, info, thr):
    """Independent formulation: exclusive cumprod per ray in float64 + autograd.
    -> (leaf sigmas f64, weights f64, exclusive transmittance f64)"""
    sig = sigmas.double().requires_grad_(True)
    outs, Ts = [], []
    for s, n in info.tolist():
        if n == 0:
            continue
        a = torch.exp(-sig[s:s + n] * steps[s:s + n].double())
        T = torch.cat([torch.ones(1, dtype=torch.float64), torch.cumprod(a, 0)[:-1]])
        alive = T.detach() > thr
        outs.append(torch.where(alive, T * (1 - a), torch.zeros_like(a)))
        Ts.append(T.detach())
    return sig, torch.cat(outs), torch.cat(Ts)


def test_weights_oracle_against_float64_autograd():
    """The reference has no CPU weights op and no test of it: pin the C restatement of
    src/cuda.cu:14-28,44-56 against an independent float64 cumprod + autograd formulation.
    Tolerance: 1e-5 relative plus T_k * 2^-22 absolute -- the reference rounds alpha to fp32 before
    forming (1 - alpha), so its weights carry an absolute error of T_k * ulp(alpha)."""
    sigmas, info, g = synthetic.packed_rays(4096, seed=3, mean_len=24, max_len=128)
    sigmas = sigmas * 8.0  # opaque enough that many rays terminate early
    steps = torch.full_like(sigmas, 5.196 / 256)
    for thr in (1e-4, 0.0):
        w = orc.weights_fwd(sigmas, steps, info, thr)
        sig64, w64, T64 = _weights_autograd_f64(sigmas, steps, info, thr)
        flips = (w == 0) != (w64 == 0)
        assert flips.sum() <= 2  # T within rounding of thr
        ok = ~flips
        err = (w.double() - w64.detach()).abs()
        assert bool((err[ok] <= 1e-5 * w64.detach().abs()[ok] + T64[ok] * 2.0 ** -22).all())
        if thr == 1e-4:
            assert (w == 0).sum() > 50, "termination not exercised"
        # backward: the reference does NOT mask terminated samples (src/cuda.cu:49-56) -> compare with thr=0 autograd
        gs = orc.weights_bwd(sigmas, steps, info, w, g)
        if thr == 0.0:
            (w64 * g.double()).sum().backward()
            scale = sig64.grad.abs().max()
            assert (gs.double() - sig64.grad).abs().max() <= 2e-5 * scale


def test_weights_oracle_edge_cases():
    # empty rays, a single-sample ray, zero samples, threshold >= 1 (nothing written)
    info = torch.tensor([[0, 0], [0, 1], [1, 0], [1, 3], [4, 0]], dtype=torch.int32)
    sig = torch.tensor([1.0, 2.0, 3.0, 4.0])
    st = torch.full((4,), 0.5)
    w = orc.weights_fwd(sig, st, info, 1e-4)
    a = torch.exp(-sig * st)
    want = torch.stack([1 - a[0], 1 - a[1], a[1] * (1 - a[2]), a[1] * a[2] * (1 - a[3])])
    assert torch.allclose(w, want, rtol=1e-6)
    assert torch.equal(orc.weights_fwd(sig, st, info, 1.0), torch.zeros(4))
    assert orc.weights_fwd(sig[:0], st[:0], info[:1], 1e-4).numel() == 0
