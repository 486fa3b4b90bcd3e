"""One training iteration of the packed-ray path as a straight sequence of C-ABI calls (no autograd graph, no
PyTorch glue kernels): the caller-side fast path of `run.Trainer` for K-Planes + the vanilla heads.

It evaluates exactly what `NerfRenderer.forward` + `MSELoss` + `loss_tv` + `backward()` evaluate through the
module/autograd path (src/core.py:225-267, src/run.py:251-259) with the same kernels in the same order; what it
removes is host work: ~70 torch ops per iteration (allocations, `cat`, strided adds, nine gradient accumulations,
the loss arithmetic) and the autograd engine's thread hop.  `tests/test_gpu_fused.py` checks it against the
autograd path on the same batch.

    fs = FusedKPlanesStep(renderer, tv_alpha=1e-4, grad_scale=2**10)
    out = fs.forward_backward(packed, info, target_rgb)      # p.grad of every parameter is set
    optimizer.step()
"""
from __future__ import annotations

import ctypes as C
import os
from typing import Dict, List

import torch
import torch.distributed as dist

from . import _cuda, _lib
from .core import NerfRenderer, is_trusted_partition, tagged_steps
from .models import (KPlanesFeatureField, VanillaColorDecoder, VanillaOpacityDecoder, _channels_last_storage,
                     _ensure_channels_last_)


def _pad4(n: int) -> int:
    return (n + 3) // 4 * 4


class FusedKPlanesStep:
    @staticmethod
    def supported(renderer: NerfRenderer) -> bool:
        fm, sd, cd = renderer.feature_module, renderer.sigma_decoder, renderer.rgb_decoder
        if not (isinstance(fm, KPlanesFeatureField) and isinstance(sd, VanillaOpacityDecoder)
                and isinstance(cd, VanillaColorDecoder)):
            return False
        if fm.dropout.p != 0.0 or not renderer.dense_rgb:
            return False
        sl, cl = sd.net.linears(), cd.net.linears()
        if len(sl) != 2 or sl[-1].out_features != 1 or cl[-1].out_features != 3:
            return False
        if any(l.out_features not in (32, 64, 128) for l in sl[:-1] + cl[:-1]):
            return False
        return all(p.is_cuda and p.dtype == torch.float32 for p in renderer.parameters())

    def __init__(self, renderer: NerfRenderer, tv_alpha: float = 0.0, grad_scale: float = 1.0, world: int = 1,
                 threshold: float = 1e-4, rank: int = 0, peer_update: bool = False):
        """peer_update (world > 1): parameters and gradients live in symmetric memory (dp.PeerMemory) and the optimiser
        update is the one-kernel reduce + Adam + broadcast over NVLink peer memory, issued from forward_backward."""
        if not self.supported(renderer):
            raise RuntimeError("FusedKPlanesStep needs KPlanesFeatureField + VanillaOpacityDecoder + VanillaColorDecoder on CUDA")
        _lib.load()
        self.renderer = renderer
        self.tv_alpha, self.grad_scale, self.world, self.threshold = float(tv_alpha), float(grad_scale), int(world), threshold
        fm: KPlanesFeatureField = renderer.feature_module  # type: ignore
        self.planes = fm._plane_params()
        self.channels = fm.plane_channels
        self.n_scales = len(self.planes) // 3
        self.res = [int(self.planes[3 * s].shape[-1]) for s in range(self.n_scales)]
        self.feat = self.n_scales * self.channels
        self.sig_lin = renderer.sigma_decoder.net.linears()   # type: ignore
        self.col_lin = renderer.rgb_decoder.net.linears()     # type: ignore
        self.n_freqs = int(renderer.rgb_decoder.pe.freqs.numel())  # type: ignore
        self.dev = self.planes[0].device
        self.bg = None if renderer.bg_color is None else (C.c_float * 3)(*[float(v) for v in renderer.bg_color.reshape(-1).tolist()])
        self.xc_width = 6 * self.n_freqs + 3 + self.feat
        self.xc_ld = _pad4(self.xc_width)
        assert self.col_lin[0].in_features == self.xc_width and self.sig_lin[0].in_features == self.feat
        # ---- one flat gradient buffer; p.grad are views of it (planes keep their channels-last strides) ----
        params: List[torch.nn.Parameter] = list(self.planes)
        for l in self.sig_lin + self.col_lin:
            params += [l.weight, l.bias]
        offs, tot = [], 0
        for p in params:
            offs.append(tot)
            tot += _pad4(p.numel())
        self.peer = None
        if peer_update and world > 1:
            from .dp import PeerMemory
            self.peer = PeerMemory(tot, self.dev, rank, world)
            self.flat_grad = self.peer.grad[:tot]
            # the parameters move into the symmetric buffer: every nn.Parameter becomes a view of it (same shape, strides
            # and values; state_dict() is unchanged), at the offset its gradient has in the gradient buffer
            with torch.no_grad():
                for p, off in zip(params, offs):
                    view = torch.as_strided(self.peer.param, p.shape, p.stride(), self.peer.param.storage_offset() + off)
                    view.copy_(p.data)
                    p.data = view
            self.peer_overlap = os.environ.get("TNF_DP_OVERLAP", "1") != "0"   # the updates on a side stream (0: in line)
            self._peer_stream = torch.cuda.Stream(device=self.dev) if self.peer_overlap else None
            # events of the side stream: the planes' / the heads' update of the latest iteration has landed on every rank.
            # The NEXT reader of the parameters on the main stream waits for them (wait_updates), not the iteration's end.
            self._planes_done = self._heads_done = None
        else:
            self.flat_grad = torch.zeros(tot, device=self.dev)
        self.grads = [torch.as_strided(self.flat_grad, p.shape, p.stride(), self.flat_grad.storage_offset() + off)
                      for p, off in zip(params, offs)]
        self._plane_grad_end = offs[len(self.planes)]   # flat_grad[:end] = plane gradients, [end:] = the heads' 
        self._scale_off = [offs[3 * sc] for sc in range(self.n_scales)] + [self._plane_grad_end]   # planes are stored scale by scale
        self._scale_bytes = [4 * (self._scale_off[sc + 1] - self._scale_off[sc]) for sc in range(self.n_scales)]
        # per-scale scatter + per-scale all-reduce (finest scale first) was measured at 2 GPUs and is slower than one scatter +
        # one all-reduce under the weight gradients (294 vs 308 M samples/s: three launches, each disturbed by the collective)
        self.bucket_scales = os.environ.get("TNF_BUCKET_SCALES", "0") != "0"
        self.params = params
        self._g = {id(p): g for p, g in zip(params, self.grads)}
        self.attach_grads()
        others = [p for p in renderer.parameters() if all(p is not q for q in params)]
        if others:
            raise RuntimeError("renderer has parameters outside the fused step")
        n = len(self.planes)
        self._plane_ptrs = (C.c_void_p * n)(*[_channels_last_storage(p).data_ptr() for p in self.planes])
        self._grad_ptrs = (C.c_void_p * n)(*[_channels_last_storage(self._g[id(p)]).data_ptr() for p in self.planes])
        self._res_scales = (C.c_int32 * self.n_scales)(*self.res)
        self._res_planes = (C.c_int32 * n)(*[int(p.shape[-1]) for p in self.planes])
        self._tv_w = (C.c_float * n)(*[1.0 / n] * n)
        self._plane_bytes = 4 * sum(p.numel() for p in self.planes)
        self._tv_gscale = torch.full((1,), self.tv_alpha / self.world * self.grad_scale, device=self.dev)
        self._tv_sums = torch.zeros(2 * n, dtype=torch.float64, device=self.dev)
        # loss_tv = mean_i (sums[2i] + sums[2i+1]) / (C (res-1) res); reported with weight tv_alpha / world
        self._tv_coef = torch.tensor([self.tv_alpha / self.world / n / float(self.channels * (r - 1) * r)
                                      for r in self._res_planes for _ in range(2)], dtype=torch.float64, device=self.dev)
        self._cap_n = self._cap_r = 0
        self._ws: Dict[str, torch.Tensor] = {}
        self.fused_composite = os.environ.get("TNF_FUSED_COMPOSITE", "1") != "0"   # 0: the three separate kernels
        self.wgrad_multi = os.environ.get("TNF_WGRAD_MULTI", "1") != "0"            # 0: one launch per 64-output weight gradient
        # Kernels of an iteration that depend on nothing before them run on an auxiliary stream beside kernels bound by a
        # different resource (TNF_AUX_OVERLAP=0: everything in line): the TV pass (HBM streaming) and the [PE(d)|d] rows
        # (issue-bound sincos) beside the plane gather (L2 -> SM bound); the colour head's output-layer backward (HBM) beside
        # the density branch's weights backward + output-layer backward.
        # Sorted scatter of the scales below the finest (csrc/kplanes.cu): the batch is counting-sorted per plane orientation
        # when it is packed (sort_batch, on the prefetch stream) and the coarse scales' plane gradients are reduced run by
        # run instead of sample by sample.  Same gradients (tests/test_gpu_fused.py) but MEASURED SLOWER than the direct
        # scatter (profiles/r02_kplanes_sorted.txt: 320 vs 270 us + a 75 us sort), so it is opt-in: TNF_KPLANES_SORTED=1.
        self.sorted_scales = (self.n_scales - 1) if (os.environ.get("TNF_KPLANES_SORTED", "0") != "0" and self.channels == 32
                                                     and self.n_scales >= 2) else 0
        self.sort_res = self.res[self.sorted_scales - 1] if self.sorted_scales else 0
        # Single GPU only by default: in data-parallel runs the update kernels already share the SMs with the step, and the
        # extra concurrency costs more than it hides (8 GPUs: 1.355 ms without, 1.40-1.45 ms with; 1 GPU: 1.303 -> 1.268 ms).
        self.aux_overlap = os.environ.get("TNF_AUX_OVERLAP", "1" if world == 1 else "0") != "0"
        self._aux = torch.cuda.Stream(device=self.dev, priority=int(os.environ.get("TNF_AUX_PRIO", "0"))) if self.aux_overlap else None
        self._aux_tv_first = os.environ.get("TNF_AUX_ORDER", "color_first") == "tv_first"
        self._closs_scratch = torch.zeros(2, dtype=torch.float64, device=self.dev)   # tnf_composite_loss_fwd_bwd (zero between calls)
        # SMs a concurrent collective kernel occupies (NCCL_MAX_CTAS when set, else the 32 CTAs NCCL uses on NVLink here:
        # measured at 2 GPUs, capping NCCL at 16 / 8 CTAs slows the step 316 -> 288 -> 244 M samples/s)
        self._sms = torch.cuda.get_device_properties(self.dev).multi_processor_count
        self.collective_ctas = int(os.environ.get("TNF_COLLECTIVE_CTAS", os.environ.get("NCCL_MAX_CTAS", "32"))) if world > 1 else 0
        # both heads' forward in one kernel (tnf_heads_fwd) when the shapes are the reference's: hidden width 64,
        # three hidden colour layers, one hidden density layer.  TNF_FUSED_HEADS=0 keeps the per-layer kernels.
        self.fused_heads = (os.environ.get("TNF_FUSED_HEADS", "1") != "0" and len(self.col_lin) == 5
                            and all(l.out_features == 64 for l in self.col_lin[:-1]) and self.sig_lin[0].out_features == 64)
        if self.fused_heads:
            tab = lambda ts: (C.c_void_p * len(ts))(*[t.data_ptr() for t in ts])
            self._cw, self._cb = tab([l.weight for l in self.col_lin]), tab([l.bias for l in self.col_lin])
            self._sw, self._sb = tab([l.weight for l in self.sig_lin]), tab([l.bias for l in self.sig_lin])
            nbytes = int(_lib.load().tnf_heads_workspace_bytes(self.feat, self.xc_width))
            self._heads_ws = torch.empty(nbytes // 4, device=self.dev)
            # ... and their data-gradient chain in one kernel (tnf_heads_bwd_data); TNF_FUSED_HEADS_BWD=0 keeps the per-layer path
            self.fused_heads_bwd = os.environ.get("TNF_FUSED_HEADS_BWD", "1") != "0" and self.feat % 32 == 0 and self.feat <= 128
            if self.fused_heads_bwd:
                self._heads_bwd_ws = torch.empty(int(_lib.load().tnf_heads_bwd_workspace_bytes(self.feat)) // 4, device=self.dev)
        else:
            self.fused_heads_bwd = False
        # With both fused head kernels the colour-input row cat([PE(d), d], features) (src/models.py:87) is never written:
        # `xc` holds only its first pe_width columns, the heads' forward and the first layer's weight gradient read the
        # rest from the feature rows (TNF_SPLIT_XC=0 keeps the materialised row).
        self.pe_width = 6 * self.n_freqs + 3
        self.split_xc = self.fused_heads and self.fused_heads_bwd and os.environ.get("TNF_SPLIT_XC", "1") != "0"
        if self.split_xc:
            self.xc_ld = _pad4(self.pe_width)
            # partial-sum tile of the first colour layer's weight gradient (zero between calls)
            self._wcat_scratch = torch.zeros(int(_lib.load().tnf_wgrad_cat_scratch_bytes(self.pe_width, self.feat)) // 4, device=self.dev)

    def attach_grads(self) -> None:
        """p.grad = its view of the flat gradient buffer (undoes optimizer.zero_grad(set_to_none=True))."""
        for p, g in zip(self.params, self.grads):
            p.grad = g

    def wait_updates(self, planes: bool = True, heads: bool = True) -> None:
        """Peer-update mode: make the current stream wait for the side stream's parameter updates of the latest iteration
        (no-op otherwise).  Every reader of the parameters calls it: the next iteration, density(), render()."""
        if self.peer is None:
            return
        cur = torch.cuda.current_stream(self.dev)
        if planes and self._planes_done is not None:
            cur.wait_event(self._planes_done)
            self._planes_done = None
        if heads and self._heads_done is not None:
            cur.wait_event(self._heads_done)
            self._heads_done = None

    # ---- per-batch sort for the sorted scatter -------------------------------------------------------
    @torch.no_grad()
    def sort_batch(self, packed: torch.Tensor) -> None:
        """Counting-sort the batch per plane orientation (tnf_kplanes_sort) on the current stream and tag `packed` with the
        result; forward_backward uses the sorted scatter when the tag is present and current."""
        n = packed.size(0)
        if not self.sorted_scales or n == 0:
            return
        cap = (n + 16383) & ~16383    # quantised like the packed rows (allocator reuse)
        pos = torch.empty(3 * cap, dtype=torch.int32, device=self.dev)[:3 * n]
        uv = torch.empty(3 * cap * 2, device=self.dev)[:3 * n * 2]
        words = int(_lib.load().tnf_kplanes_sort_scratch_ints(self.sort_res, cap))
        scratch = torch.empty(words, dtype=torch.int32, device=self.dev)
        with torch.cuda.device(self.dev):
            _lib.call("tnf_kplanes_sort", packed.data_ptr(), 7, n, self.sort_res, scratch.data_ptr(), pos.data_ptr(), uv.data_ptr(),
                      _lib.stream_ptr(), nbytes=n * (12 + 3 * 16), extra_kernels=3)
        packed._tnf_ksort = (pos, uv, packed._version)

    @staticmethod
    def sorted_tag(packed: torch.Tensor):
        tag = getattr(packed, "_tnf_ksort", None)
        return tag if tag is not None and tag[2] == packed._version else None

    # ---- workspace ---------------------------------------------------------------------------------
    def _reserve(self, n: int, r: int) -> None:
        if n > self._cap_n:
            cap = int(n * 1.25) + 1024
            hid_c = self.col_lin[0].out_features
            hid_s = self.sig_lin[0].out_features
            e = lambda *s: torch.empty(*s, device=self.dev)
            ws = self._ws
            ws["feats"], ws["dfeat"] = e(cap, self.feat), e(cap, self.feat)
            ws["hs"], ws["dhs"] = e(cap, hid_s), e(cap, hid_s)
            ws["sigma"], ws["gsigma"], ws["w"], ws["gw"] = e(cap), e(cap), e(cap), e(cap)
            ws["xc"] = e(cap, self.xc_ld)
            if not self.fused_heads_bwd:
                ws["dxc"] = e(cap, self.xc_ld)
            for i in range(len(self.col_lin) - 1):
                ws[f"h{i}"] = e(cap, hid_c)
            for i in range(len(self.col_lin) - 1):
                ws[f"dh{i}"] = e(cap, hid_c)
            ws["rgb"], ws["grgb"] = e(cap, 3), e(cap, 3)
            if self.sorted_scales:
                ws["krows"] = e(3 * self.sorted_scales * cap * self.channels)   # per-plane gradient rows in slot order
            self._cap_n = cap
        if r > self._cap_r:
            cap = int(r * 1.25) + 256
            self._ws["rendered"] = torch.empty(cap, 3, device=self.dev)
            self._ws["grend"] = torch.empty(cap, 3, device=self.dev)
            self._ws["loss"] = torch.zeros(1, device=self.dev)
            self._cap_r = cap

    # ---- density only (the occupancy update's sigma_fn, src/run.py:249) ----------------------------
    @torch.no_grad()
    def density(self, coords: torch.Tensor) -> torch.Tensor:
        """sigma_decoder(feature_module(coords)) for [n,3] contracted coordinates: the same two kernels the modules run,
        on the iteration's own workspaces (no allocations, no torch glue).  Returns a view of the workspace that stays
        valid until the next density() / forward_backward() call on this stream."""
        _lib.require_cuda(coords, "coords")
        if coords.dim() != 2 or coords.size(1) != 3 or coords.dtype != torch.float32 or not coords.is_contiguous():
            raise RuntimeError("coords must be a contiguous [n,3] float32 tensor")
        n = coords.size(0)
        if n == 0:
            return torch.empty(0, 1, device=self.dev)
        self._reserve(n, 1)
        self.wait_updates()
        ws, call, st = self._ws, _lib.call, _lib.stream_ptr()
        F, l0, l1 = self.feat, self.sig_lin[0], self.sig_lin[1]
        with torch.cuda.device(self.dev):
            call("tnf_kplanes_fwd", self._plane_ptrs, self._res_scales, self.n_scales, self.channels, coords.data_ptr(), 3, n,
                 ws["feats"].data_ptr(), st, nbytes=n * (12 + 4 * F) + self._plane_bytes)
            call("tnf_linear_fwd", ws["feats"].data_ptr(), F, l0.weight.data_ptr(), l0.bias.data_ptr(), None, l0.out_features, n,
                 l0.out_features, F, 1, l1.weight.data_ptr(), l1.bias.data_ptr(), ws["sigma"].data_ptr(), 1, 1, st,
                 nbytes=4 * (n * (F + 1) + l0.out_features * F), flops=2 * n * l0.out_features * (F + 1))
        return ws["sigma"][:n].view(n, 1)

    # ---- forward only: the render half of the path (src/run.py:34-42, NerfRenderer.forward in eval mode) -----------
    @torch.no_grad()
    def render(self, packed: torch.Tensor, info: torch.Tensor, out: torch.Tensor | None = None) -> torch.Tensor:
        """rendered [R,3] of a packed batch: gather -> both heads -> weights -> composite, five launches, nothing saved for
        a backward pass (tnf_heads_fwd is called without activation outputs).  `out` [R,3] is an optional destination."""
        _lib.require_cuda(packed, "packed_samples")
        n, r = packed.size(0), info.size(0)
        if not (self.fused_heads and self.split_xc):
            raise RuntimeError("FusedKPlanesStep.render needs the fused heads kernel (reference head shapes)")
        if not (packed.is_contiguous() and info.is_contiguous() and info.dtype == torch.int32):
            raise RuntimeError("packed samples / packing info must be contiguous ([N,7] fp32, [R,2] int32)")
        if out is None:
            out = torch.empty(r, 3, device=self.dev)
        if n == 0 or r == 0:
            raise ValueError("no samples remaining")
        steps = tagged_steps(packed)
        sstride = 1
        if steps is None:
            steps, sstride = packed[:, 6], 7
        flags = _cuda.TRUSTED_PARTITION if is_trusted_partition(info) else 0
        status = None if flags else torch.empty(1, dtype=torch.int32, device=self.dev)
        self._reserve_render(n)
        self.wait_updates()
        ws, call, st = self._rws, _lib.call, _lib.stream_ptr()
        P = lambda t: t.data_ptr()
        F, xw, xld = self.feat, self.xc_width, self.xc_ld
        with torch.cuda.device(self.dev):
            call("tnf_kplanes_fwd", self._plane_ptrs, self._res_scales, self.n_scales, self.channels, P(packed), 7, n,
                 P(ws["feats"]), st, nbytes=n * (12 + 4 * F) + self._plane_bytes)
            call("tnf_color_input", P(packed) + 12, 7, P(ws["feats"]), F, self.n_freqs, 0, P(ws["xc"]), xld, n, st,
                 nbytes=n * (12 + 4 * xld))
            call("tnf_heads_fwd", P(ws["feats"]), F, F, P(ws["xc"]), xld, xw, self.pe_width, self._cw, self._cb, self._sw,
                 self._sb, None, None, P(ws["rgb"]), P(ws["sigma"]), n, P(self._heads_ws), st,
                 nbytes=4 * n * (F + xld + 4), flops=2 * n * (64 * (F + 1) + 64 * xw + 3 * 64 * 64 + 3 * 64))
            call("tnf_weights_fwd", P(ws["sigma"]), P(steps), sstride, P(info), float(self.threshold), P(ws["w"]), n, r,
                 flags, _lib.ptr(status), st, nbytes=12 * n + 8 * r, extra_kernels=0 if flags else 3)
            call("tnf_composite_fwd", P(ws["w"]), P(ws["rgb"]), P(info), n, r, self.bg, P(out), None, st,
                 nbytes=16 * n + 20 * r)
        return out

    def _reserve_render(self, n: int) -> None:
        """Forward-only workspaces (feature rows, [PE(d)|d] rows, sigma, rgb, weights): 4*(96+52+5) B per sample."""
        if n > getattr(self, "_rcap", 0):
            cap = (n + 65535) & ~65535
            e = lambda *s: torch.empty(*s, device=self.dev)
            self._rws = {"feats": e(cap, self.feat), "xc": e(cap, self.xc_ld), "sigma": e(cap), "rgb": e(cap, 3), "w": e(cap)}
            self._rcap = cap

    # ---- the iteration -----------------------------------------------------------------------------
    @torch.no_grad()
    def forward_backward(self, packed: torch.Tensor, info: torch.Tensor, target: torch.Tensor,
                         n_rays_global: torch.Tensor | None = None, reduce: bool = False, n_rays_work=None,
                         after_plane_grads=None, peer_step: dict | None = None) -> Dict[str, torch.Tensor]:
        """packed [N,7], info [R,2] int32 (a RayProvider partition), target [R,3].  Sets p.grad of every parameter to
        d(grad_scale * (MSE_union + tv_alpha/world * loss_tv))/dp and returns {"loss", "rendered"}.  With `reduce` the
        gradients are all-reduced over the ranks (sum) before returning, overlapped with the tail of backward.
        peer_step = {"step", "lr", "betas", "eps", "weight_decay"} (peer_update mode): the union batch's ray count comes
        from the peers' slot tables and the optimiser update of iteration `step` (1-based) happens in here -- the planes'
        right after their gradients are complete, the heads' after the weight gradients, on a side stream.  On return the
        updates are enqueued, not joined: whoever reads the parameters next on another stream calls wait_updates() first (the
        next forward_backward, density() and render() do); p.grad holds this rank's un-reduced gradients."""
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
        ws, call, st = self._ws, _lib.call, _lib.stream_ptr()
        P = lambda t: t.data_ptr()
        G = lambda p: self._g[id(p)].data_ptr()
        F, xw, xld = self.feat, self.xc_width, self.xc_ld
        sl, cl = self.sig_lin, self.col_lin
        hs_w, hc_w = sl[0].out_features, cl[0].out_features
        nh = len(cl) - 1  # hidden layers of the colour head

        def lin_fwd(x, ldx, lin, y, head=None, head_out=None, head_act=0):
            nn_, k = lin.out_features, lin.in_features
            n_head = 0 if head is None else head.out_features
            call("tnf_linear_fwd", x, ldx, P(lin.weight), P(lin.bias), P(y), nn_, n, nn_, k, 1,
                 None if head is None else P(head.weight), None if head is None else P(head.bias),
                 None if head is None else P(head_out), n_head, head_act, st,
                 nbytes=4 * (n * (k + nn_ + n_head) + nn_ * k), flops=2 * n * nn_ * (k + n_head))

        def wgrad(dy, lddy, x, ldx, lin):
            nn_, k = lin.out_features, lin.in_features
            call("tnf_linear_bwd_weight", dy, lddy, x, ldx, G(lin.weight), G(lin.bias), n, nn_, k, st,
                 nbytes=4 * (n * (nn_ + k) + nn_ * k), flops=2 * n * nn_ * k)

        def dgrad(dy, lddy, lin, dx, lddx, relu_src, ldrs):
            nn_, k = lin.out_features, lin.in_features
            call("tnf_linear_bwd_data", dy, lddy, P(lin.weight), dx, lddx, relu_src, ldrs, n, nn_, k, st,
                 nbytes=4 * (n * (nn_ + (2 if relu_src else 1) * k) + nn_ * k), flops=2 * n * nn_ * k)

        with torch.cuda.device(self.dev):
            # gradient buffer: the TV pass WRITES the plane gradients (value + gradient of the regulariser from one read of
            # the planes; loss += tv_alpha * loss_tv, src/run.py:254-255) -- it depends on nothing in this iteration, so it
            # doubles as the zero-fill of 99.8 % of the buffer; the data term is scattered on top of it in backward
            defer_heads = peer_step is not None and self.peer_overlap and self.tv_alpha != 0.0
            self.wait_updates(planes=True, heads=not defer_heads)   # the planes are read first (TV pass, gather); the heads later
            aux_on = self.aux_overlap and self.fused_heads and self.split_xc and self.tv_alpha != 0.0
            main = torch.cuda.current_stream(self.dev)
            xc_ready = tv_done = None
            if aux_on:
                aux = self._aux
                aux.wait_event(main.record_event())   # everything of the previous iteration (and the waits above) first
                with torch.cuda.stream(aux):
                    ast = aux.cuda_stream
                    for which in (("tv", "color") if self._aux_tv_first else ("color", "tv")):
                        if which == "color":
                            call("tnf_color_input", P(packed) + 12, 7, P(ws["feats"]), F, self.n_freqs, 0, P(ws["xc"]), xld, n, ast,
                                 nbytes=n * (12 + 4 * xld))
                            xc_ready = aux.record_event()
                        else:
                            call("tnf_tv_fwd_bwd", self._plane_ptrs, self._grad_ptrs, self._res_planes, len(self.planes),
                                 self.channels, self._tv_w, P(self._tv_gscale), 0, P(self._tv_sums), ast, nbytes=2 * self._plane_bytes)
                            tv_done = aux.record_event()
                for t_ in (packed, target):   # read on the auxiliary stream too
                    t_.record_stream(aux)
                if not defer_heads:
                    self.flat_grad[self._plane_grad_end:].zero_()
            elif self.tv_alpha != 0.0:
                call("tnf_tv_fwd_bwd", self._plane_ptrs, self._grad_ptrs, self._res_planes, len(self.planes), self.channels,
                     self._tv_w, P(self._tv_gscale), 0, P(self._tv_sums), st, nbytes=2 * self._plane_bytes)
                if not defer_heads:
                    self.flat_grad[self._plane_grad_end:].zero_()
            else:
                self.flat_grad.zero_()
            # ---- forward (src/core.py:225-267) ----
            call("tnf_kplanes_fwd", self._plane_ptrs, self._res_scales, self.n_scales, self.channels, P(packed), 7, n,
                 P(ws["feats"]), st, nbytes=n * (12 + 4 * F) + self._plane_bytes)
            if defer_heads:
                # the heads' update of the previous iteration (side stream) must have landed before their weights are read,
                # and every rank must have consumed the heads' gradients before they are zeroed for this iteration
                self.wait_updates(planes=False, heads=True)
                self.flat_grad[self._plane_grad_end:].zero_()
            if self.fused_heads:
                xc_feat = 0 if self.split_xc else F   # feature columns copied into the colour-input row
                if xc_ready is not None:
                    main.wait_event(xc_ready)
                else:
                    call("tnf_color_input", P(packed) + 12, 7, P(ws["feats"]), F, self.n_freqs, xc_feat, P(ws["xc"]), xld, n, st,
                         nbytes=n * (12 + 4 * xc_feat + 4 * xld))
                hptrs = (C.c_void_p * 4)(*[P(ws[f"h{i}"]) for i in range(4)])
                mlp_flops = 2 * n * (64 * (F + 1) + 64 * xw + 3 * 64 * 64 + 3 * 64)
                call("tnf_heads_fwd", P(ws["feats"]), F, F, P(ws["xc"]), xld, xw, self.pe_width if self.split_xc else xw,
                     self._cw, self._cb, self._sw, self._sb, hptrs,
                     P(ws["hs"]), P(ws["rgb"]), P(ws["sigma"]), n, P(self._heads_ws), st,
                     nbytes=4 * n * (F + xld + 5 * 64 + 4), flops=mlp_flops)
            else:
                lin_fwd(P(ws["feats"]), F, sl[0], ws["hs"], head=sl[1], head_out=ws["sigma"], head_act=1)
            call("tnf_weights_fwd", P(ws["sigma"]), P(steps), sstride, P(info), float(self.threshold), P(ws["w"]), n, r,
                 flags, _lib.ptr(status), st, nbytes=12 * n + 8 * r, extra_kernels=0 if flags else 3)
            if not self.fused_heads:
                call("tnf_color_input", P(packed) + 12, 7, P(ws["feats"]), F, self.n_freqs, F, P(ws["xc"]), xld, n, st,
                     nbytes=n * (12 + 4 * F + 4 * xld))
            x, ldx = P(ws["xc"]), xld
            for i in range(0 if self.fused_heads else nh):
                last = i == nh - 1
                lin_fwd(x, ldx, cl[i], ws[f"h{i}"], head=cl[-1] if last else None,
                        head_out=ws["rgb"] if last else None, head_act=2 if last else 0)
                x, ldx = P(ws[f"h{i}"]), hc_w
            # ---- composite + loss + their gradients in one pass over the rays (src/core.py:256-265, src/run.py:252,259) ----
            if n_rays_work is not None:
                n_rays_work.wait()   # the union batch's ray count (async all-reduce started before the forward)
            if peer_step is not None:
                n_rays_global = self.peer.sum_counts(peer_step["step"])   # published by every rank at the start of its step
            tv_in_loss = self.fused_composite and self.tv_alpha != 0.0   # the kernel adds tv_alpha/world * loss_tv to the loss it reports
            if tv_done is not None:
                main.wait_event(tv_done)   # the loss kernel reads the regulariser's sums; the scatter adds onto its gradients
            if self.fused_composite:
                call("tnf_composite_loss_fwd_bwd", P(ws["w"]), P(ws["rgb"]), P(info), n, r, self.bg, P(target), float(r),
                     _lib.ptr(n_rays_global), self.grad_scale, P(ws["rendered"]), P(ws["gw"]), P(ws["grgb"]), P(ws["loss"]),
                     P(self._closs_scratch), P(self._tv_sums) if tv_in_loss else None, P(self._tv_coef) if tv_in_loss else None,
                     self._tv_sums.numel() if tv_in_loss else 0, st, nbytes=48 * n + 56 * r)
            else:
                call("tnf_composite_fwd", P(ws["w"]), P(ws["rgb"]), P(info), n, r, self.bg, P(ws["rendered"]), None, st,
                     nbytes=16 * n + 20 * r)
                call("tnf_mse_loss_grad", P(ws["rendered"]), P(target), r, float(r), _lib.ptr(n_rays_global), self.grad_scale,
                     P(ws["grend"]), P(ws["loss"]), st, nbytes=36 * r)
                call("tnf_composite_bwd", P(ws["w"]), P(ws["rgb"]), P(info), n, r, self.bg, P(ws["grend"]), P(ws["gw"]),
                     P(ws["grgb"]), st, nbytes=32 * n + 20 * r)
            # ---- backward ----
            # Order: the data-gradient chain first (it ends in the plane gradients, 99.8 % of the bytes a data-parallel run
            # has to all-reduce), then the weight gradients of the heads -- 0.5 ms of work that depends only on the saved
            # dh / activations and hides the plane all-reduce completely.
            # colour head: fused output layer + sigmoid, then the hidden layers' data gradients from the last to the first
            dh = [ws[f"dh{i}"] for i in range(nh)]   # dh[i] = gradient wrt the pre-activation of colour layer i
            colour_done = None
            if aux_on and self.fused_heads_bwd:
                aux.wait_event(main.record_event())
                with torch.cuda.stream(aux):
                    call("tnf_head_bwd", P(ws[f"h{nh - 1}"]), hc_w, P(cl[-1].weight), P(ws["rgb"]), P(ws["grgb"]), P(dh[nh - 1]),
                         G(cl[-1].weight), G(cl[-1].bias), n, hc_w, 3, 2, aux.cuda_stream, nbytes=4 * n * (2 * hc_w + 6))
                    colour_done = aux.record_event()
            else:
                call("tnf_head_bwd", P(ws[f"h{nh - 1}"]), hc_w, P(cl[-1].weight), P(ws["rgb"]), P(ws["grgb"]), P(dh[nh - 1]),
                     G(cl[-1].weight), G(cl[-1].bias), n, hc_w, 3, 2, st, nbytes=4 * n * (2 * hc_w + 6))
            if not self.fused_heads_bwd:
                for i in range(nh - 1, 0, -1):
                    inp = ws[f"h{i - 1}"]
                    dgrad(P(dh[i]), hc_w, cl[i], P(dh[i - 1]), hc_w, P(inp), hc_w)
                dgrad(P(dh[0]), hc_w, cl[0], P(ws["dxc"]), xld, None, 0)
            # density branch: weights backward, fused output layer + truncated_exp, hidden layer
            call("tnf_weights_bwd", P(ws["sigma"]), P(steps), sstride, P(info), P(ws["w"]), P(ws["gw"]), P(ws["gsigma"]), n, r,
                 flags, _lib.ptr(status), st, nbytes=20 * n + 8 * r, extra_kernels=0 if flags else 3)
            call("tnf_head_bwd", P(ws["hs"]), hs_w, P(sl[1].weight), P(ws["sigma"]), P(ws["gsigma"]), P(ws["dhs"]),
                 G(sl[1].weight), G(sl[1].bias), n, hs_w, 1, 1, st, nbytes=4 * n * (2 * hs_w + 2))
            dfeat = ws["dfeat"][:n]
            if colour_done is not None:
                main.wait_event(colour_done)
            if self.fused_heads_bwd:
                # the whole data-gradient chain of both heads, down to the feature rows, in one kernel
                masks = (C.c_void_p * 3)(*[P(ws[f"h{i}"]) for i in (2, 1, 0)])
                dh_out = (C.c_void_p * 3)(*[P(dh[i]) for i in (2, 1, 0)])
                call("tnf_heads_bwd_data", P(dh[3]), P(ws["dhs"]), masks, self._cw, xw, xw - F, P(sl[0].weight), F, dh_out,
                     P(ws["dfeat"]), F, n, P(self._heads_bwd_ws), st, nbytes=4 * n * (2 * 64 + 3 * 64 + 3 * 64 + F),
                     flops=2 * n * 64 * (3 * 64 + 2 * F))
            else:
                dgrad(P(ws["dhs"]), hs_w, sl[0], P(ws["dfeat"]), F, None, 0)
                # the features feed both heads: d feats = d(sigma branch) + d(colour input)[:, feature columns]
                torch.add(dfeat, ws["dxc"][:n, xw - F:xw], out=dfeat)
                _lib.launch_count += 1
            # plane gradients: scatter-add of the data term on top of the TV gradient written at the start
            work, works = None, []
            ksort = self.sorted_tag(packed) if self.sorted_scales else None
            kp_sorted = lambda phase: call(
                "tnf_kplanes_bwd_sorted", self._plane_ptrs, self._grad_ptrs, self._res_scales, self.n_scales, self.channels, P(packed), 7, n,
                P(dfeat), self.sorted_scales, P(ksort[0]), P(ksort[1]), P(ws["krows"]), phase, st,
                label="tnf_kplanes_bwd" + ("" if phase == 0 else f"(phase {phase})"), extra_kernels=1 if phase == 0 else 0,
                nbytes=(n * (12 + 4 * F) + 2 * self._plane_bytes) if phase != 2 else 0)
            if reduce and self.world > 1 and self.bucket_scales:
                # data-parallel: finest scale first (76 % of the bytes); its all-reduce starts while the coarser scales are
                # still being scattered and then runs on under the heads' weight gradients
                for sc in range(self.n_scales - 1, -1, -1):
                    call("tnf_kplanes_bwd_scales", self._plane_ptrs, self._grad_ptrs, self._res_scales, self.n_scales, self.channels,
                         P(packed), 7, n, P(dfeat), sc, sc + 1, st, nbytes=(n * (12 + 4 * F)) // self.n_scales + 2 * self._scale_bytes[sc])
                    works.append(dist.all_reduce(self.flat_grad[self._scale_off[sc]:self._scale_off[sc + 1]], async_op=True))
                work = works.pop()
            elif peer_step is not None:
                # Data-parallel update over NVLink peer memory: gradient sum over the ranks + Adam on this rank's share + new
                # planes to every rank, one kernel per range (99.9 % of the parameter bytes; nothing below depends on it).  The
                # kernel is NVLink-bound (~280 us for the 132 MB of planes whatever the rank count), so it runs on a side stream:
                # the finest scale (76 % of the bytes) is scattered first and its update starts while the coarser scales are
                # still being scattered, then runs on under the heads' weight gradients.
                ps = peer_step
                upd = lambda lo, hi, st_: self.peer.reduce_adam_bcast(lo, hi, 0, ps["step"], ps["lr"], ps["betas"], ps["eps"],
                                                                      ps["weight_decay"], st_)
                kp_bwd = lambda a, b: call("tnf_kplanes_bwd_scales", self._plane_ptrs, self._grad_ptrs, self._res_scales, self.n_scales,
                                           self.channels, P(packed), 7, n, P(dfeat), a, b, st, label="tnf_kplanes_bwd",
                                           nbytes=((n * (12 + 4 * F)) * (b - a)) // self.n_scales + 2 * sum(self._scale_bytes[a:b]))
                if self.peer_overlap:
                    main, comm = torch.cuda.current_stream(self.dev), self._peer_stream
                    fine = self.n_scales - 1
                    if ksort is not None and self.sorted_scales == fine:
                        # phase 1 finishes the finest scale (direct scatter) and stores the coarse scales' rows; the finest
                        # scale's update then runs beside phase 2 (the run-merging scatter of the coarse scales)
                        kp_sorted(1)
                        comm.wait_event(main.record_event())
                        with torch.cuda.stream(comm):
                            upd(self._scale_off[fine], self._scale_off[self.n_scales], comm.cuda_stream)
                        kp_sorted(2)
                        comm.wait_event(main.record_event())
                        with torch.cuda.stream(comm):
                            upd(0, self._scale_off[fine], comm.cuda_stream)
                    else:
                        for a, b in ((fine, self.n_scales), (0, fine)):
                            if a == b:
                                continue
                            kp_bwd(a, b)
                            comm.wait_event(main.record_event())
                            with torch.cuda.stream(comm):
                                upd(self._scale_off[a], self._scale_off[b], comm.cuda_stream)
                    planes_done = comm.record_event()
                else:
                    kp_sorted(0) if ksort is not None else kp_bwd(0, self.n_scales)
                    upd(0, self._plane_grad_end, st)
            else:
                if ksort is not None:
                    kp_sorted(0)
                else:
                    call("tnf_kplanes_bwd", self._plane_ptrs, self._grad_ptrs, self._res_scales, self.n_scales, self.channels,
                         P(packed), 7, n, P(dfeat), st, nbytes=n * (12 + 4 * F) + 2 * self._plane_bytes)
                if reduce and self.world > 1:
                    work = dist.all_reduce(self.flat_grad[:self._plane_grad_end], async_op=True)  # runs under the wgrads
            if after_plane_grads is not None and work is None:
                after_plane_grads()   # the plane gradients are final (single-GPU): their optimiser update may start now
            if work is not None and self.collective_ctas > 0:
                # the collective's CTAs hold whole SMs; leave them out of the weight-gradient kernels' one-CTA-per-SM grids
                _lib.load().tnf_set_sm_budget(max(16, self._sms - self.collective_ctas))
            # weight gradients of both heads
            multi = self.wgrad_multi and hc_w == 64 and hs_w == 64 and F <= 128 and 1 <= nh <= 4
            if multi:
                # the hidden colour layers' and the density layer's weight gradients as jobs of ONE launch (one cold start and
                # one drain instead of nh: ~17 us of every 43 us launch, profiles/r01_role_timing_wgrad.txt)
                jl = [(dh[i], ws[f"h{i - 1}"], hc_w, cl[i]) for i in range(nh - 1, 0, -1)] + [(ws["dhs"], ws["feats"], F, sl[0])]
                nj = len(jl)
                vp, i64, i32 = (C.c_void_p * nj), (C.c_int64 * nj), (C.c_int32 * nj)
                call("tnf_linear_bwd_weight_multi", nj, vp(*[P(j[0]) for j in jl]), i64(*[64] * nj), vp(*[P(j[1]) for j in jl]),
                     i64(*[j[2] for j in jl]), i32(*[j[2] for j in jl]), vp(*[G(j[3].weight) for j in jl]),
                     vp(*[G(j[3].bias) for j in jl]), n, st,
                     nbytes=sum(4 * (n * (64 + j[2]) + 64 * j[2]) for j in jl), flops=sum(2 * n * 64 * j[2] for j in jl))
            else:
                for i in range(nh - 1, 0, -1):
                    wgrad(P(dh[i]), hc_w, P(ws[f"h{i - 1}"]), hc_w, cl[i])
            if self.split_xc:
                call("tnf_linear_bwd_weight_cat", P(dh[0]), hc_w, P(ws["xc"]), xld, self.pe_width, P(ws["feats"]), F, F,
                     G(cl[0].weight), G(cl[0].bias), n, hc_w, P(self._wcat_scratch), st, nbytes=4 * (n * (hc_w + xld + F) + hc_w * xw),
                     flops=2 * n * hc_w * xw)
            else:
                wgrad(P(dh[0]), hc_w, P(ws["xc"]), xld, cl[0])
            if not multi:
                wgrad(P(ws["dhs"]), hs_w, P(ws["feats"]), F, sl[0])
            if peer_step is not None:
                # the heads' 28 K parameters the same way, now that their gradients are complete
                hupd = lambda st_: self.peer.reduce_adam_bcast(self._plane_grad_end, self.peer.n, 1, peer_step["step"], peer_step["lr"],
                                                               peer_step["betas"], peer_step["eps"], peer_step["weight_decay"], st_)
                if self.peer_overlap:
                    # on the side stream too: the main stream goes on to the next iteration's TV pass and gather, which read
                    # only the planes; the heads' kernel (two rank barriers: ~40-90 us of latency and skew) leaves its path
                    comm.wait_event(main.record_event())
                    with torch.cuda.stream(comm):
                        hupd(comm.cuda_stream)
                        self._heads_done = comm.record_event()
                    self._planes_done = planes_done
                else:
                    hupd(st)
            if work is not None:
                _lib.load().tnf_set_sm_budget(0)
                dist.all_reduce(self.flat_grad[self._plane_grad_end:])
                for w in works:
                    w.wait()
                work.wait()
            loss = ws["loss"][0]
            if self.tv_alpha != 0.0 and not tv_in_loss:
                loss = loss + torch.dot(self._tv_sums, self._tv_coef).float()
            else:
                loss = loss.clone()
        return {"loss": loss, "rendered": ws["rendered"][:r]}
