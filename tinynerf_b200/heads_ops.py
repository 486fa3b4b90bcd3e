"""Both decoder heads of the module/autograd path in the fused tcgen05 kernels.

`NerfRenderer.forward` (src/core.py:225-267) evaluates `sigma_decoder(features)` and `rgb_decoder(features, dirs)`
(VanillaOpacityDecoder / VanillaColorDecoder, src/models.py:70-89).  When both have the reference's shapes (hidden width 64,
one hidden density layer, three hidden colour layers, feature width a multiple of 32 up to 128: K-Planes 96, Cobafa 128)
`fused_heads` evaluates them together: forward = tnf_color_input + tnf_heads_fwd (one persistent kernel for both heads of a
128-sample tile, activations kept in tensor memory), backward = tnf_head_bwd x2 + tnf_heads_bwd_data (the whole data-gradient
chain in one kernel) + the TMA weight-gradient kernels -- the same entry points fused.FusedKPlanesStep strings together,
here behind torch.autograd so that any feature module (Cobafa with its trunk, K-Planes through the modules) can sit in front.
Other shapes (the vanilla model's 256-wide features) keep the per-layer kernels of mlp_ops.
"""
from __future__ import annotations

import ctypes as C
from typing import Any

import torch
from torch.autograd import Function

from . import _lib


def supported(sigma_decoder, rgb_decoder, features: torch.Tensor) -> bool:
    from .models import VanillaColorDecoder, VanillaOpacityDecoder
    if not (isinstance(sigma_decoder, VanillaOpacityDecoder) and isinstance(rgb_decoder, VanillaColorDecoder)):
        return False
    if not (features.is_cuda and features.dtype == torch.float32 and features.dim() == 2):
        return False
    sl, cl = sigma_decoder.net.linears(), rgb_decoder.net.linears()
    F = features.size(1)
    if len(sl) != 2 or len(cl) != 5 or sl[1].out_features != 1 or cl[4].out_features != 3:
        return False
    if any(l.out_features != 64 for l in cl[:4]) or sl[0].out_features != 64 or sl[0].in_features != F:
        return False
    n_freqs = int(rgb_decoder.pe.freqs.numel())
    if F % 32 != 0 or not (32 <= F <= 128) or cl[0].in_features != 6 * n_freqs + 3 + F or n_freqs > 24:
        return False
    return all(p.is_cuda and p.dtype == torch.float32 for l in sl + cl for p in (l.weight, l.bias))


def _pad4(n: int) -> int:
    return (n + 3) // 4 * 4


class _FusedHeads(Function):
    @staticmethod
    def forward(ctx: Any, feats: torch.Tensor, dirs: torch.Tensor, n_freqs: int, *params: torch.Tensor):  # type: ignore
        """params = (cw0, cb0, ..., cw4, cb4, sw0, sb0, sw1, sb1) -> (sigma [n,1], rgb [n,3])"""
        lib = _lib.load()
        cw, cb, sw, sb = params[0:10:2], params[1:10:2], params[10:14:2], params[11:14:2]
        feats = feats.contiguous()
        n, F = feats.shape
        dev = feats.device
        pe_w = 6 * n_freqs + 3
        k0, xld = pe_w + F, _pad4(pe_w)
        if dirs.dim() != 2 or dirs.size(1) != 3 or dirs.stride(1) != 1 or dirs.dtype != torch.float32:
            dirs = dirs.reshape(-1, 3).float().contiguous()
        save = any(ctx.needs_input_grad)
        e = lambda *s: torch.empty(*s, device=dev)
        xc, rgb, sigma = e(n, xld), e(n, 3), e(n)
        hs = [e(n, 64) for _ in range(4)] if save else None
        hsig = e(n, 64) if save else None
        tab = lambda ts: (C.c_void_p * len(ts))(*[t.data_ptr() for t in ts])
        ws = torch.empty(int(lib.tnf_heads_workspace_bytes(F, k0)) // 4, device=dev)
        if n > 0:
            with torch.cuda.device(dev):
                st = _lib.stream_ptr()
                _lib.call("tnf_color_input", dirs.data_ptr(), dirs.stride(0), feats.data_ptr(), F, n_freqs, 0, xc.data_ptr(), xld, n, st,
                          nbytes=n * (12 + 4 * xld))
                _lib.call("tnf_heads_fwd", feats.data_ptr(), F, F, xc.data_ptr(), xld, k0, pe_w, tab(cw), tab(cb), tab(sw), tab(sb),
                          tab(hs) if save else None, _lib.ptr(hsig), rgb.data_ptr(), sigma.data_ptr(), n, ws.data_ptr(), st,
                          nbytes=4 * n * (F + xld + (5 * 64 if save else 0) + 4),
                          flops=2 * n * (64 * (F + 1) + 64 * k0 + 3 * 64 * 64 + 3 * 64))
        if save:
            ctx.save_for_backward(feats, xc, rgb, sigma, hsig, *hs, *params)
            ctx.dims = (n, F, pe_w, k0, xld)
        return sigma.view(n, 1), rgb

    @staticmethod
    def backward(ctx: Any, g_sigma: torch.Tensor, g_rgb: torch.Tensor):  # type: ignore
        lib = _lib.load()
        feats, xc, rgb, sigma, hsig, h0, h1, h2, h3 = ctx.saved_tensors[:9]
        params = ctx.saved_tensors[9:]
        cw, sw = params[0:10:2], params[10:14:2]
        n, F, pe_w, k0, xld = ctx.dims
        dev = feats.device
        zeros = lambda t: torch.zeros_like(t)
        gcw, gcb = [zeros(w) for w in cw], [zeros(b) for b in params[1:10:2]]
        gsw, gsb = [zeros(w) for w in sw], [zeros(b) for b in params[11:14:2]]
        dfeat = torch.empty(n, F, device=dev)
        if n > 0:
            g_sigma = (torch.zeros(n, device=dev) if g_sigma is None else g_sigma.reshape(-1).float().contiguous())
            g_rgb = (torch.zeros(n, 3, device=dev) if g_rgb is None else g_rgb.float().contiguous())
            e = lambda *s: torch.empty(*s, device=dev)
            dh = [e(n, 64) for _ in range(4)]
            dhs = e(n, 64)
            P = lambda t: t.data_ptr()
            call = _lib.call
            with torch.cuda.device(dev):
                st = _lib.stream_ptr()
                call("tnf_head_bwd", P(h3), 64, P(cw[4]), P(rgb), P(g_rgb), P(dh[3]), P(gcw[4]), P(gcb[4]), n, 64, 3, 2, st,
                     nbytes=4 * n * (2 * 64 + 6))
                call("tnf_head_bwd", P(hsig), 64, P(sw[1]), P(sigma), P(g_sigma), P(dhs), P(gsw[1]), P(gsb[1]), n, 64, 1, 1, st,
                     nbytes=4 * n * (2 * 64 + 2))
                masks = (C.c_void_p * 3)(P(h2), P(h1), P(h0))
                dh_out = (C.c_void_p * 3)(P(dh[2]), P(dh[1]), P(dh[0]))
                cwt = (C.c_void_p * 5)(*[P(w) for w in cw])
                wsb = torch.empty(int(lib.tnf_heads_bwd_workspace_bytes(F)) // 4, device=dev)
                call("tnf_heads_bwd_data", P(dh[3]), P(dhs), masks, cwt, k0, k0 - F, P(sw[0]), F, dh_out, P(dfeat), F, n, P(wsb), st,
                     nbytes=4 * n * (2 * 64 + 3 * 64 + 3 * 64 + F), flops=2 * n * 64 * (3 * 64 + 2 * F))
                acts = [h0, h1, h2]
                for i in (3, 2, 1):   # dW_i += dh_i^T h_{i-1}
                    call("tnf_linear_bwd_weight", P(dh[i]), 64, P(acts[i - 1]), 64, P(gcw[i]), P(gcb[i]), n, 64, 64, st,
                         nbytes=4 * (n * 128 + 64 * 64), flops=2 * n * 64 * 64)
                # first colour layer: input row = [PE(d) | d] (xc) ++ feature row; the TMA kernel takes five 32-column atoms,
                # so a 128-wide feature row goes in two launches (columns [0, 96) with the PE part, then [96, 128))
                fa = min(F, 32 * (5 - (pe_w + 31) // 32))
                if fa == F:
                    scratch = torch.zeros(int(lib.tnf_wgrad_cat_scratch_bytes(pe_w, F)) // 4, device=dev)
                    call("tnf_linear_bwd_weight_cat", P(dh[0]), 64, P(xc), xld, pe_w, P(feats), F, F, P(gcw[0]), P(gcb[0]), n, 64,
                         P(scratch), st, nbytes=4 * (n * (64 + xld + F) + 64 * k0), flops=2 * n * 64 * k0)
                else:
                    part_a = torch.zeros(64, pe_w + fa, device=dev)
                    part_b = torch.zeros(64, F - fa, device=dev)
                    scratch = torch.zeros(int(lib.tnf_wgrad_cat_scratch_bytes(pe_w, fa)) // 4, device=dev)
                    call("tnf_linear_bwd_weight_cat", P(dh[0]), 64, P(xc), xld, pe_w, P(feats), F, fa, P(part_a), P(gcb[0]), n, 64,
                         P(scratch), st, nbytes=4 * (n * (64 + xld + fa) + 64 * (pe_w + fa)), flops=2 * n * 64 * (pe_w + fa))
                    call("tnf_linear_bwd_weight", P(dh[0]), 64, P(feats) + 4 * fa, F, P(part_b), None, n, 64, F - fa, st,
                         nbytes=4 * (n * (64 + F - fa) + 64 * (F - fa)), flops=2 * n * 64 * (F - fa))
                    gcw[0][:, :pe_w + fa] = part_a
                    gcw[0][:, pe_w + fa:] = part_b
                call("tnf_linear_bwd_weight", P(dhs), 64, P(feats), F, P(gsw[0]), P(gsb[0]), n, 64, F, st,
                     nbytes=4 * (n * (64 + F) + 64 * F), flops=2 * n * 64 * F)
        else:
            dfeat.zero_()
        grads = []
        for w, b in zip(gcw, gcb):
            grads += [w, b]
        for w, b in zip(gsw, gsb):
            grads += [w, b]
        return (dfeat, None, None, *grads)


def fused_heads(sigma_decoder, rgb_decoder, features: torch.Tensor, dirs: torch.Tensor):
    """-> (sigmas [n,1], rgbs [n,3]) == (sigma_decoder(features), rgb_decoder(features, dirs))"""
    sl, cl = sigma_decoder.net.linears(), rgb_decoder.net.linears()
    params = []
    for l in cl:
        params += [l.weight, l.bias]
    for l in sl:
        params += [l.weight, l.bias]
    return _FusedHeads.apply(features, dirs, int(rgb_decoder.pe.freqs.numel()), *params)
