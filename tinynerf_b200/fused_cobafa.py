"""One Cobafa training iteration of the packed-ray path as a straight sequence of C-ABI calls (BASELINE config 3): the
counterpart of fused.FusedKPlanesStep for `CobafaFeatureField` + the vanilla heads (src/run.py:141-150).

It evaluates what `NerfRenderer.forward` + `MSELoss` + `backward()` evaluate through the module/autograd path
(src/core.py:225-267, src/models.py:258-266, src/run.py:251-259) with the same kernels in the same order:

    tnf_cobafa_fwd -> Dropout(0.01) -> 7 x tnf_linear_fwd (the 36 -> 128 -> ... -> 128 trunk) -> tnf_color_input ->
    tnf_heads_fwd -> tnf_weights_fwd -> tnf_composite_loss_fwd_bwd -> tnf_head_bwd x2 -> tnf_weights_bwd ->
    tnf_heads_bwd_data -> the heads' weight gradients -> 7 x (tnf_linear_bwd_weight, tnf_linear_bwd_data) ->
    dropout backward -> tnf_cobafa_bwd

on caller-owned workspaces with ONE flat gradient buffer (p.grad are views of it: one zero-fill, and one collective in a
data-parallel run).  What it removes is host work: the module path spends ~4 ms of Python / autograd / allocator time on an
iteration whose kernels take 3 ms (scripts/host_profile.py cobafa), i.e. the GPU waits for the host.  The dropout mask comes
from `torch.native_dropout` on the same [N, 36] rows -- the kernel and the generator stream `torch.nn.Dropout` uses in the
module path, so both paths draw the same mask from the same seed (tests/test_gpu_fused.py).
"""
from __future__ import annotations

import ctypes as C
import os
from typing import Dict, List

import torch
import torch.distributed as dist

from . import _cuda, _lib
from .core import NerfRenderer, is_trusted_partition, tagged_steps
from .models import (CobafaFeatureField, VanillaColorDecoder, VanillaOpacityDecoder, _channels_last_storage,
                     _ensure_channels_last_)


def _pad4(n: int) -> int:
    return (n + 3) // 4 * 4


class FusedCobafaStep:
    @staticmethod
    def supported(renderer: NerfRenderer) -> bool:
        fm, sd, cd = renderer.feature_module, renderer.sigma_decoder, renderer.rgb_decoder
        if not (isinstance(fm, CobafaFeatureField) and isinstance(sd, VanillaOpacityDecoder) and isinstance(cd, VanillaColorDecoder)):
            return False
        if not renderer.dense_rgb or not fm.mlp._relu:
            return False
        tl, sl, cl = fm.mlp.linears(), sd.net.linears(), cd.net.linears()
        F = fm.feature_dim
        if len(tl) < 2 or tl[-1].out_features != F or F % 32 != 0 or not (32 <= F <= 128):
            return False
        if any(l.out_features % 32 != 0 or l.out_features > 128 or l.in_features > 160 or l.bias is None for l in tl):
            return False
        if any(b.grid.shape[2] != b.grid.shape[3] or b.grid.shape[3] != b.grid.shape[4] for b in [fm.coef_grid, *fm.basis_grids]):
            return False
        if len(sl) != 2 or len(cl) != 5 or sl[1].out_features != 1 or cl[4].out_features != 3:
            return False
        if any(l.out_features != 64 for l in cl[:4]) or sl[0].out_features != 64 or sl[0].in_features != F:
            return False
        n_freqs = int(cd.pe.freqs.numel())
        if cl[0].in_features != 6 * n_freqs + 3 + F or n_freqs > 24:
            return False
        return all(p.is_cuda and p.dtype == torch.float32 for p in renderer.parameters())

    def __init__(self, renderer: NerfRenderer, grad_scale: float = 1.0, world: int = 1, threshold: float = 1e-4):
        if not self.supported(renderer):
            raise RuntimeError("FusedCobafaStep needs CobafaFeatureField + VanillaOpacityDecoder + VanillaColorDecoder on CUDA")
        lib = _lib.load()
        self.renderer = renderer
        self.grad_scale, self.world, self.threshold = float(grad_scale), int(world), threshold
        fm: CobafaFeatureField = renderer.feature_module  # type: ignore
        self.fm = fm
        self.grids = [fm.coef_grid.grid, *[g.grid for g in fm.basis_grids]]   # type: ignore
        for g in self.grids:
            _ensure_channels_last_(g)
        self.trunk = fm.mlp.linears()
        self.sig_lin = renderer.sigma_decoder.net.linears()   # type: ignore
        self.col_lin = renderer.rgb_decoder.net.linears()     # type: ignore
        self.n_freqs = int(renderer.rgb_decoder.pe.freqs.numel())  # type: ignore
        self.feat = int(fm.feature_dim)
        self.dev = self.grids[0].device
        self.bg = None if renderer.bg_color is None else (C.c_float * 3)(*[float(v) for v in renderer.bg_color.reshape(-1).tolist()])
        self.pe_width = 6 * self.n_freqs + 3
        self.xc_width = self.pe_width + self.feat
        self.xc_ld = _pad4(self.pe_width)
        # ---- lookup tables of the fused gather (models._CobafaLookup) ----
        basis = self.grids[1:]
        L = len(basis)
        self.n_levels = L
        self._res = (C.c_int32 * L)(*[int(b.shape[-1]) for b in basis])
        self._ch = (C.c_int32 * L)(*[int(b.shape[1]) for b in basis])
        self._freqs = (C.c_float * L)(*[float(enc.f) for enc in fm.encoders])   # type: ignore
        self.feat_in = sum(int(b.shape[1]) for b in basis)     # width of the concatenated basis*coef rows (36)
        self.coef_res = int(self.grids[0].shape[-1])
        assert self.trunk[0].in_features == self.feat_in
        # ---- one flat gradient buffer; p.grad are views of it (the grids keep their channels-last strides) ----
        params: List[torch.nn.Parameter] = list(self.grids)
        for l in self.trunk + self.sig_lin + self.col_lin:
            params += [l.weight, l.bias]
        offs, tot = [], 0
        for p in params:
            offs.append(tot)
            tot += _pad4(p.numel())
        self.flat_grad = torch.zeros(tot, device=self.dev)
        self.grads = [torch.as_strided(self.flat_grad, p.shape, p.stride(), self.flat_grad.storage_offset() + off)
                      for p, off in zip(params, offs)]
        self.params = params
        self._g = {id(p): g for p, g in zip(params, self.grads)}
        self.attach_grads()
        if any(all(p is not q for q in params) for p in renderer.parameters()):
            raise RuntimeError("renderer has parameters outside the fused step")
        self._basis_ptrs = (C.c_void_p * L)(*[_channels_last_storage(b).data_ptr() for b in basis])
        self._gbasis_ptrs = (C.c_void_p * L)(*[_channels_last_storage(self._g[id(b)]).data_ptr() for b in basis])
        self._coef_ptr = _channels_last_storage(self.grids[0]).data_ptr()
        self._gcoef_ptr = _channels_last_storage(self._g[id(self.grids[0])]).data_ptr()
        self._grid_bytes = 4 * sum(g.numel() for g in self.grids)
        tab = lambda ts: (C.c_void_p * len(ts))(*[t.data_ptr() for t in ts])
        self._cw, self._cb = tab([l.weight for l in self.col_lin]), tab([l.bias for l in self.col_lin])
        self._sw, self._sb = tab([l.weight for l in self.sig_lin]), tab([l.bias for l in self.sig_lin])
        self._heads_ws = torch.empty(int(lib.tnf_heads_workspace_bytes(self.feat, self.xc_width)) // 4, device=self.dev)
        self._heads_bwd_ws = torch.empty(int(lib.tnf_heads_bwd_workspace_bytes(self.feat)) // 4, device=self.dev)
        # first colour layer's weight gradient: the TMA kernel takes five 32-column atoms of [PE(d) | d | features], a
        # 128-wide feature row goes in two launches (heads_ops._FusedHeads.backward)
        self._fa = min(self.feat, 32 * (5 - (self.pe_width + 31) // 32))
        self._wcat_scratch = torch.zeros(int(lib.tnf_wgrad_cat_scratch_bytes(self.pe_width, self._fa)) // 4, device=self.dev)
        self._closs_scratch = torch.zeros(2, dtype=torch.float64, device=self.dev)
        self._cap_n = self._cap_r = 0
        self._ws: Dict[str, torch.Tensor] = {}
        self.wgrad_multi = os.environ.get("TNF_WGRAD_MULTI", "1") != "0"   # 0: one launch per 64-output weight gradient

    def attach_grads(self) -> None:
        """p.grad = its view of the flat gradient buffer (undoes optimizer.zero_grad(set_to_none=True))."""
        for p, g in zip(self.params, self.grads):
            p.grad = g

    def _reserve(self, n: int, r: int) -> None:
        if n > self._cap_n:
            cap = int(n * 1.25) + 1024
            e = lambda *s: torch.empty(*s, device=self.dev)
            ws = self._ws
            ws["f36"], ws["df36"] = e(cap, self.feat_in), e(cap, self.feat_in)
            for i, l in enumerate(self.trunk[:-1]):
                ws[f"t{i}"] = e(cap, l.out_features)              # activations of the trunk's hidden layers
            wmax = max(l.out_features for l in self.trunk)
            ws["dta"], ws["dtb"] = e(cap, wmax), e(cap, wmax)     # the trunk's data gradients, ping-pong
            ws["feats"], ws["dfeat"] = e(cap, self.feat), e(cap, self.feat)
            ws["hs"], ws["dhs"] = e(cap, 64), e(cap, 64)
            ws["sigma"], ws["gsigma"], ws["w"], ws["gw"] = e(cap), e(cap), e(cap), e(cap)
            ws["xc"] = e(cap, self.xc_ld)
            for i in range(4):
                ws[f"h{i}"], ws[f"dh{i}"] = e(cap, 64), e(cap, 64)
            ws["rgb"], ws["grgb"] = e(cap, 3), e(cap, 3)
            self._cap_n = cap
        if r > self._cap_r:
            cap = int(r * 1.25) + 256
            self._ws["rendered"] = torch.empty(cap, 3, device=self.dev)
            self._ws["loss"] = torch.zeros(1, device=self.dev)
            self._cap_r = cap

    # ---- density only (the occupancy update's sigma_fn, src/run.py:249) ----------------------------
    @torch.no_grad()
    def density(self, coords: torch.Tensor) -> torch.Tensor:
        """sigma_decoder(feature_module(coords)) for [n,3] contracted coordinates on the iteration's own workspaces: lookup,
        dropout when the module is in training mode (the reference updates the grid from inside the training loop, so its
        Dropout(0.01) is active there too), trunk, density head with the fused output layer.  Returns a view of the workspace
        that stays valid until the next density() / forward_backward() call on this stream."""
        _lib.require_cuda(coords, "coords")
        if coords.dim() != 2 or coords.size(1) != 3 or coords.dtype != torch.float32 or not coords.is_contiguous():
            raise RuntimeError("coords must be a contiguous [n,3] float32 tensor")
        n = coords.size(0)
        if n == 0:
            return torch.empty(0, 1, device=self.dev)
        self._reserve(n, 1)
        ws, call, st = self._ws, _lib.call, _lib.stream_ptr(self.dev)
        P = lambda t: t.data_ptr()
        F, F0, tl, sl = self.feat, self.feat_in, self.trunk, self.sig_lin
        p_drop = float(self.fm.dropout.p) if self.fm.dropout.training else 0.0
        with torch.cuda.device(self.dev):
            call("tnf_cobafa_fwd", self._basis_ptrs, self._res, self._ch, self._freqs, self.n_levels, self._coef_ptr, self.coef_res,
                 P(coords), 3, n, P(ws["f36"]), st, nbytes=n * (12 + 4 * F0) + self._grid_bytes)
            x0 = ws["f36"][:n]
            if p_drop > 0.0:
                x0 = torch.native_dropout(x0, p_drop, True)[0]
            x, ldx = P(x0), x0.stride(0)
            for i, lin in enumerate(tl):
                last = i == len(tl) - 1
                y = ws["feats"] if last else ws[f"t{i}"]
                nn_, k = lin.out_features, lin.in_features
                call("tnf_linear_fwd", x, ldx, P(lin.weight), P(lin.bias), P(y), nn_, n, nn_, k, int(not last), None, None, None, 0, 0, st,
                     nbytes=4 * (n * (k + nn_) + nn_ * k), flops=2 * n * nn_ * k)
                x, ldx = P(y), nn_
            l0, l1 = sl[0], sl[1]
            call("tnf_linear_fwd", P(ws["feats"]), F, P(l0.weight), P(l0.bias), None, l0.out_features, n, l0.out_features, F, 1,
                 P(l1.weight), P(l1.bias), P(ws["sigma"]), 1, 1, st,
                 nbytes=4 * (n * (F + 1) + l0.out_features * F), flops=2 * n * l0.out_features * (F + 1))
        return ws["sigma"][:n].view(n, 1)

    @torch.no_grad()
    def forward_backward(self, packed: torch.Tensor, info: torch.Tensor, target: torch.Tensor,
                         n_rays_global: torch.Tensor | None = None, reduce: bool = False, n_rays_work=None) -> Dict[str, torch.Tensor]:
        """packed [N,7], info [R,2] int32 (a RayProvider partition), target [R,3].  Sets p.grad of every parameter to
        d(grad_scale * MSE_union)/dp and returns {"loss", "rendered"}.  With `reduce` the flat gradient is all-reduced over the
        ranks (sum, one collective) before returning."""
        _lib.require_cuda(packed, "packed_samples")
        n, r = packed.size(0), info.size(0)
        if n == 0 or r == 0:
            raise ValueError("no samples remaining")
        if not (packed.is_contiguous() and info.is_contiguous() and info.dtype == torch.int32):
            raise RuntimeError("packed samples / packing info must be contiguous ([N,7] fp32, [R,2] int32)")
        target = target.contiguous()
        steps = tagged_steps(packed)
        sstride = 1
        if steps is None:
            steps, sstride = packed[:, 6], 7
        flags = _cuda.TRUSTED_PARTITION if is_trusted_partition(info) else 0
        status = None if flags else torch.empty(1, dtype=torch.int32, device=self.dev)
        self._reserve(n, r)
        self.attach_grads()
        ws, call, st = self._ws, _lib.call, _lib.stream_ptr(self.dev)
        P = lambda t: t.data_ptr()
        G = lambda p: self._g[id(p)].data_ptr()
        F, F0, xw, xld, pe_w = self.feat, self.feat_in, self.xc_width, self.xc_ld, self.pe_width
        tl, sl, cl = self.trunk, self.sig_lin, self.col_lin
        p_drop = float(self.fm.dropout.p) if self.fm.dropout.training else 0.0

        def lin_fwd(x, ldx, lin, y, relu):
            nn_, k = lin.out_features, lin.in_features
            call("tnf_linear_fwd", x, ldx, P(lin.weight), P(lin.bias), P(y), nn_, n, nn_, k, int(relu), None, None, None, 0, 0, st,
                 nbytes=4 * (n * (k + nn_) + nn_ * k), flops=2 * n * nn_ * k)

        def wgrad(dy, lddy, x, ldx, lin):
            nn_, k = lin.out_features, lin.in_features
            call("tnf_linear_bwd_weight", dy, lddy, x, ldx, G(lin.weight), G(lin.bias), n, nn_, k, st,
                 nbytes=4 * (n * (nn_ + k) + nn_ * k), flops=2 * n * nn_ * k)

        def dgrad(dy, lddy, lin, dx, lddx, relu_src, ldrs):
            nn_, k = lin.out_features, lin.in_features
            call("tnf_linear_bwd_data", dy, lddy, P(lin.weight), dx, lddx, relu_src, ldrs, n, nn_, k, st,
                 nbytes=4 * (n * (nn_ + (2 if relu_src else 1) * k) + nn_ * k), flops=2 * n * nn_ * k)

        with torch.cuda.device(self.dev):
            self.flat_grad.zero_()   # the grid gradients are scattered with reductions, the weight gradients accumulate
            # ---- forward (src/models.py:258-266, src/core.py:225-267) ----
            call("tnf_cobafa_fwd", self._basis_ptrs, self._res, self._ch, self._freqs, self.n_levels, self._coef_ptr, self.coef_res,
                 P(packed), 7, n, P(ws["f36"]), st, nbytes=n * (12 + 4 * F0) + self._grid_bytes)
            f36 = ws["f36"][:n]
            drop_mask = None
            if p_drop > 0.0:
                x0, drop_mask = torch.native_dropout(f36, p_drop, True)   # torch.nn.Dropout's kernel and generator stream
            else:
                x0 = f36
            x, ldx = P(x0), x0.stride(0)
            for i, lin in enumerate(tl):
                last = i == len(tl) - 1
                y = ws["feats"] if last else ws[f"t{i}"]
                lin_fwd(x, ldx, lin, y, relu=not last)
                x, ldx = P(y), lin.out_features
            call("tnf_color_input", P(packed) + 12, 7, P(ws["feats"]), F, self.n_freqs, 0, P(ws["xc"]), xld, n, st,
                 nbytes=n * (12 + 4 * xld))
            hptrs = (C.c_void_p * 4)(*[P(ws[f"h{i}"]) for i in range(4)])
            call("tnf_heads_fwd", P(ws["feats"]), F, F, P(ws["xc"]), xld, xw, pe_w, self._cw, self._cb, self._sw, self._sb, hptrs,
                 P(ws["hs"]), P(ws["rgb"]), P(ws["sigma"]), n, P(self._heads_ws), st,
                 nbytes=4 * n * (F + xld + 5 * 64 + 4), flops=2 * n * (64 * (F + 1) + 64 * xw + 3 * 64 * 64 + 3 * 64))
            call("tnf_weights_fwd", P(ws["sigma"]), P(steps), sstride, P(info), float(self.threshold), P(ws["w"]), n, r,
                 flags, _lib.ptr(status), st, nbytes=12 * n + 8 * r, extra_kernels=0 if flags else 3)
            # ---- composite + loss + their gradients in one pass over the rays (src/core.py:256-265, src/run.py:252,259) ----
            if n_rays_work is not None:
                n_rays_work.wait()   # the union batch's ray count (async all-reduce started before the forward)
            call("tnf_composite_loss_fwd_bwd", P(ws["w"]), P(ws["rgb"]), P(info), n, r, self.bg, P(target), float(r),
                 _lib.ptr(n_rays_global), self.grad_scale, P(ws["rendered"]), P(ws["gw"]), P(ws["grgb"]), P(ws["loss"]),
                 P(self._closs_scratch), None, None, 0, st, nbytes=48 * n + 56 * r)
            # ---- backward: the heads (heads_ops._FusedHeads.backward) ----
            dh = [ws[f"dh{i}"] for i in range(4)]
            call("tnf_head_bwd", P(ws["h3"]), 64, P(cl[4].weight), P(ws["rgb"]), P(ws["grgb"]), P(dh[3]), G(cl[4].weight), G(cl[4].bias),
                 n, 64, 3, 2, st, nbytes=4 * n * (2 * 64 + 6))
            call("tnf_weights_bwd", P(ws["sigma"]), P(steps), sstride, P(info), P(ws["w"]), P(ws["gw"]), P(ws["gsigma"]), n, r,
                 flags, _lib.ptr(status), st, nbytes=20 * n + 8 * r, extra_kernels=0 if flags else 3)
            call("tnf_head_bwd", P(ws["hs"]), 64, P(sl[1].weight), P(ws["sigma"]), P(ws["gsigma"]), P(ws["dhs"]), G(sl[1].weight),
                 G(sl[1].bias), n, 64, 1, 1, st, nbytes=4 * n * (2 * 64 + 2))
            masks = (C.c_void_p * 3)(*[P(ws[f"h{i}"]) for i in (2, 1, 0)])
            dh_out = (C.c_void_p * 3)(*[P(dh[i]) for i in (2, 1, 0)])
            call("tnf_heads_bwd_data", P(dh[3]), P(ws["dhs"]), masks, self._cw, xw, xw - F, P(sl[0].weight), F, dh_out,
                 P(ws["dfeat"]), F, n, P(self._heads_bwd_ws), st, nbytes=4 * n * (2 * 64 + 3 * 64 + 3 * 64 + F),
                 flops=2 * n * 64 * (3 * 64 + 2 * F))
            # dW_i += dh_i^T h_{i-1} of the hidden colour layers and the density layer: four jobs of one launch
            jl = [(dh[i], ws[f"h{i - 1}"], 64, cl[i]) for i in (3, 2, 1)] + [(ws["dhs"], ws["feats"], F, sl[0])]
            if self.wgrad_multi:
                nj = len(jl)
                vp, i64, i32 = (C.c_void_p * nj), (C.c_int64 * nj), (C.c_int32 * nj)
                call("tnf_linear_bwd_weight_multi", nj, vp(*[P(j[0]) for j in jl]), i64(*[64] * nj), vp(*[P(j[1]) for j in jl]),
                     i64(*[j[2] for j in jl]), i32(*[j[2] for j in jl]), vp(*[G(j[3].weight) for j in jl]),
                     vp(*[G(j[3].bias) for j in jl]), n, st,
                     nbytes=sum(4 * (n * (64 + j[2]) + 64 * j[2]) for j in jl), flops=sum(2 * n * 64 * j[2] for j in jl))
            else:
                for dy_, x_, k_, lin_ in jl:
                    wgrad(P(dy_), 64, P(x_), k_, lin_)
            fa = self._fa
            if fa == F:
                call("tnf_linear_bwd_weight_cat", P(dh[0]), 64, P(ws["xc"]), xld, pe_w, P(ws["feats"]), F, F, G(cl[0].weight),
                     G(cl[0].bias), n, 64, P(self._wcat_scratch), st, nbytes=4 * (n * (64 + xld + F) + 64 * xw), flops=2 * n * 64 * xw)
            else:
                part_a = torch.zeros(64, pe_w + fa, device=self.dev)
                part_b = torch.zeros(64, F - fa, device=self.dev)
                call("tnf_linear_bwd_weight_cat", P(dh[0]), 64, P(ws["xc"]), xld, pe_w, P(ws["feats"]), F, fa, P(part_a), G(cl[0].bias),
                     n, 64, P(self._wcat_scratch), st, nbytes=4 * (n * (64 + xld + fa) + 64 * (pe_w + fa)), flops=2 * n * 64 * (pe_w + fa))
                call("tnf_linear_bwd_weight", P(dh[0]), 64, P(ws["feats"]) + 4 * fa, F, P(part_b), None, n, 64, F - fa, st,
                     nbytes=4 * (n * (64 + F - fa) + 64 * (F - fa)), flops=2 * n * 64 * (F - fa))
                gw0 = self._g[id(cl[0].weight)]
                gw0[:, :pe_w + fa] = part_a
                gw0[:, pe_w + fa:] = part_b
            # ---- backward: the trunk (mlp_ops._FusedMLP.backward, wide last layer without activation) ----
            dcur, ldc = P(ws["dfeat"]), F
            bufs = (ws["dta"], ws["dtb"])
            for i in range(len(tl) - 1, -1, -1):
                lin = tl[i]
                if i > 0:
                    inp, ldi = P(ws[f"t{i - 1}"]), tl[i - 1].out_features
                else:
                    inp, ldi = P(x0), x0.stride(0)
                wgrad(dcur, ldc, inp, ldi, lin)
                if i > 0:
                    dnext = bufs[i & 1]
                    dgrad(dcur, ldc, lin, P(dnext), lin.in_features, inp, ldi)    # masked by the ReLU that produced this layer's input
                    dcur, ldc = P(dnext), lin.in_features
                else:
                    dgrad(dcur, ldc, lin, P(ws["df36"]), F0, None, 0)
            d36 = ws["df36"][:n]
            if drop_mask is not None:
                d36 = torch.ops.aten.native_dropout_backward(d36, drop_mask, 1.0 / (1.0 - p_drop))
            call("tnf_cobafa_bwd", self._basis_ptrs, self._gbasis_ptrs, self._res, self._ch, self._freqs, self.n_levels, self._coef_ptr,
                 self._gcoef_ptr, self.coef_res, P(packed), 7, n, P(d36), st, nbytes=n * (12 + 4 * F0) + 2 * self._grid_bytes)
            if reduce and self.world > 1:
                dist.all_reduce(self.flat_grad)
            loss = ws["loss"][0].clone()
        return {"loss": loss, "rendered": ws["rendered"][:r]}
