"""Achieved worst-case errors of the CUDA path against the reference restatement.  TEST INFRASTRUCTURE ONLY.

The same measurements back two consumers: tests/test_gpu_fused.py asserts the bars on them, and bench.py prints them as
the `parity` object of its JSON line (the relaxations argued in DESIGN.md section 2 are reported as
numbers where a reader of the bench line sees them).  Everything here runs on the GPU box: oracle/ref_port.py on "cuda"
(stock grid_sample / Linear / index_add_ -- the arithmetic the reference itself executes there) with the weights op
through the UNMODIFIED reference kernel (oracle/_ref/_cuda.so).  Nothing under tinynerf_b200/ imports this file.
"""
from __future__ import annotations

import torch

from . import load_ref_cuda
from . import ref_port as rp

DEV = "cuda"


def _relu_kink(layers, x64, margin):
    near = torch.zeros(x64.size(0), dtype=torch.bool, device=x64.device)
    h = x64
    for w, b in layers[:-1]:
        pre = torch.nn.functional.linear(h, w.double(), b.double())
        near |= (pre.abs() < margin).any(1)
        h = pre.relu()
    return near


def kplanes_case(n_rays: int = 9600, seed: int = 0):
    """BASELINE config 2 shape: K-Planes + vanilla heads, AABB +-1.5, analytic 128^3 grid, ~2^18 packed samples."""
    from tinynerf_b200 import core, models, synthetic
    torch.manual_seed(seed)
    field = models.KPlanesFeatureField(32)
    sd = models.VanillaOpacityDecoder(96)
    cd = models.VanillaColorDecoder(8, 96, 64, 3)
    with torch.no_grad():
        sd.net.net[-1].bias += 3.0
    renderer = core.NerfRenderer(field, sd, cd, bg_color=torch.ones(3)).to(DEV)
    aabb = torch.tensor([[-1.5, -1.5, -1.5], [1.5, 1.5, 1.5]], device=DEV)
    marcher = core.RayMarcherAABB(aabb, 256, 0.1)
    og = core.OccupancyGrid(128, marcher.step_size, 0.01, synthetic.DECAY).to(DEV)
    og.grid.copy_(synthetic.analytic_grid(128, seed=1))
    og.mean = og.grid.mean().item()
    prov = core.RayProvider(og, core.ContractionAABB(aabb), marcher)
    o, d = synthetic.blender_rays(n_rays, seed=2)
    return renderer, prov, og, aabb, o.to(DEV), d.to(DEV)


def march_parity(n_rays: int = 4096) -> dict:
    """a4-a10: packing info and packed rows of a training batch, bit for bit against the restatement on the same GPU."""
    renderer, prov, og, aabb, o, d = kplanes_case(n_rays)
    torch.manual_seed(77)
    packed, info = prov(o, d, True)
    torch.manual_seed(77)
    noise = torch.rand(n_rays, 256, device=DEV)
    want_p, want_i, mask = rp.ray_provider(o, d, og.grid, og.threshold, scene="aabb", n_samples=256, aabb=aabb, near=0.1,
                                           far=1e5, noise=noise)
    return {"n_rays": n_rays, "n_samples": int(packed.size(0)), "kept_frac": round(float(mask.float().mean()), 4),
            "packing_info_bit_exact": bool(torch.equal(info, want_i)),
            "packed_rows_bit_exact": bool(packed.shape == want_p.shape and torch.equal(packed, want_p))}


def weights_parity(logn: int = 20, thr: float = 1e-4) -> dict:
    """a1/a2 through the trainer's TRUSTED_PARTITION path against the reference kernel on the same inputs."""
    from tinynerf_b200 import _cuda, synthetic
    ref = load_ref_cuda()
    if ref is None:
        return {"unavailable": "oracle/_ref/_cuda.so missing"}
    sig, info, g = [t.to(DEV) for t in synthetic.packed_rays(1 << logn, seed=1000 + logn)]
    sig = sig * 4.0
    steps = torch.full_like(sig, 5.196 / 256)
    w_ref = ref.compute_weights_fwd(sig, steps, info, thr)
    w = _cuda.weights_fwd(sig, steps, info, thr, _cuda.TRUSTED_PARTITION)
    live = w_ref > 0
    g_ref = ref.compute_weights_bwd(sig, steps, info, w_ref, g)
    gs = _cuda.weights_bwd(sig, steps, info, w_ref, g, _cuda.TRUSTED_PARTITION)
    # the reference forms -sum_{j>k} w_j g_j as prefix - total in fp32: entries carry an absolute error ~ step * sum_ray |w g|
    ray_id = torch.repeat_interleave(torch.arange(info.size(0), device=DEV), info[:, 1].long())
    ray_abs = torch.zeros(info.size(0), device=DEV, dtype=torch.float64).index_add_(0, ray_id, (w_ref * g).abs().double())
    bound = 1e-5 * g_ref.abs().double() + 1e-5 * steps.double() * ray_abs[ray_id]
    return {"n_samples": 1 << logn, "n_rays": int(info.size(0)), "threshold": thr,
            "termination_mask_bit_exact": bool(torch.equal(w > 0, live)),
            "weights_max_rel_err": float(((w - w_ref).abs()[live] / w_ref[live]).max()),
            "grad_sigmas_max_err_over_bound": float(((gs - g_ref).abs().double() / bound.clamp_min(1e-300)).max())}


def fused_step_parity(n_rays: int = 9600) -> dict:
    """The iteration bench.py times (fused.FusedKPlanesStep.forward_backward) directly against the restatement: rendered
    colours, loss and every parameter gradient of a ~2^18-sample batch.  Rays holding a sample whose hidden pre-activation
    lies within 3e-6 of a ReLU kink (decided differently by two correct fp32 evaluations) are removed from the batch of
    BOTH pipelines; their share is reported."""
    from tinynerf_b200.fused import FusedKPlanesStep
    renderer, prov, og, aabb, o, d = kplanes_case(n_rays)
    field, sd, cd = renderer.feature_module, renderer.sigma_decoder, renderer.rgb_decoder
    torch.manual_seed(5)
    packed, info = prov(o, d, training=True)
    s_layers = [(l.weight, l.bias) for l in sd.net.linears()]
    c_layers = [(l.weight, l.bias) for l in cd.net.linears()]
    planes = [[p.plane for p in scale] for scale in field.planes]
    with torch.no_grad():
        feats = rp.kplanes_features(planes, packed[:, :3])
        xcol = torch.cat([rp.positional_encoding(packed[:, 3:6], 8), packed[:, 3:6], feats], -1).double()
        kink = _relu_kink(s_layers, feats.double(), 3e-6) | _relu_kink(c_layers, xcol, 3e-6)
        ray_id = torch.repeat_interleave(torch.arange(info.size(0), device=DEV), info[:, 1].long())
        bad_ray = torch.zeros(info.size(0), device=DEV).index_add_(0, ray_id, kink.float()) > 0
        keep_s = ~bad_ray[ray_id]
        packed2 = packed[keep_s].contiguous()
        cnt = info[~bad_ray, 1]
        info2 = torch.stack([torch.cumsum(cnt, 0, dtype=torch.int32) - cnt, cnt], -1).contiguous()
    n, r = packed2.size(0), info2.size(0)
    target = torch.rand(r, 3, device=DEV)
    tv_alpha, gscale = 1e-4, 2.0 ** 10
    fs = FusedKPlanesStep(renderer, tv_alpha=tv_alpha, grad_scale=gscale)
    res = fs.forward_backward(packed2, info2, target)
    mine = {k: p.grad.clone() for k, p in renderer.named_parameters()}
    out, loss = res["rendered"].clone(), float(res["loss"])
    for p in renderer.parameters():
        p.grad = None
    want = rp.render(lambda x: rp.kplanes_features(planes, x), lambda f: rp.sigma_head(s_layers, f),
                     lambda f, dd: rp.rgb_head(c_layers, 8, f, dd), packed2, info2, torch.ones(3))
    ref_loss = torch.nn.functional.mse_loss(want, target) + tv_alpha * rp.kplanes_tv(planes)
    (ref_loss * gscale).backward()
    err = (out.double() - want.double()).abs()
    excess = (err - (1e-5 * want.double().abs() + 2e-6)).max().item()
    per = {}
    for k, p in renderer.named_parameters():
        ref, got = p.grad.double(), mine[k].double()
        scale = ref.abs().max().clamp_min(1e-12)
        per[k] = (((got - ref).norm() / ref.norm().clamp_min(1e-30)).item(), ((got - ref).abs().max() / scale).item())
    plane_keys = [k for k in per if "plane" in k]
    head_keys = [k for k in per if k not in plane_keys]
    return {"n_samples": int(n), "n_rays": int(r), "kink_rays_excluded_frac": round(float(bad_ray.float().mean()), 4),
            "rendered_max_abs_err": float(err.max()), "rendered_excess_over_1e-5rel+2e-6": float(excess),
            "loss_rel_err": abs(loss - float(ref_loss)) / abs(float(ref_loss)),
            "plane_grad_rel_l2_max": max(per[k][0] for k in plane_keys),
            "plane_grad_max_err_over_tensor_max": max(per[k][1] for k in plane_keys),
            "head_grad_rel_l2_max": max(per[k][0] for k in head_keys),
            "head_grad_max_err_over_tensor_max": max(per[k][1] for k in head_keys),
            "per_parameter": {k: (float(f"{a:.3e}"), float(f"{b:.3e}")) for k, (a, b) in per.items()}}
