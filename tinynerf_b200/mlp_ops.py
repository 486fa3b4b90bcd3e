"""The MLP heads on tcgen05 tensor cores: autograd wrapper over tnf_linear_fwd / _bwd_data / _bwd_weight /
tnf_head_bwd (include/tinynerf_b200.h, csrc/mlp.cu).

`fused_mlp(x, linears, head_act)` evaluates the reference's `MLP` module (src/models.py:7-28: Linear, ReLU,
..., Linear) whose LAST Linear has <= 4 outputs (sigma: 1, rgb: 3) with that last layer and its activation
fused into the epilogue of the last hidden layer:
    head_act 0: identity            (MLP.forward)
    head_act 1: truncated_exp(v-1.) (VanillaOpacityDecoder, src/models.py:70-77)
    head_act 2: sigmoid             (VanillaColorDecoder,  src/models.py:79-89)
`fused_trunk(x, linears)` evaluates an MLP whose last Linear is wide (Cobafa trunk, 128 outputs, no activation).
"""
from __future__ import annotations

from typing import List, Sequence

import torch
from torch.autograd import Function

from . import _lib

MAX_IN = 160   # resident-weight kernels (csrc/mlp.cu): in_features <= 160, out_features a multiple of 32 up to 128
_HEAD_ACT_TO_WIDE = {0: 0, 1: 2, 2: 3}   # fused-head activation code -> tnf_wide_linear_fwd `act`


def _resident(n: int, k: int) -> bool:
    """Layer [n, k] fits the resident-weight tcgen05 kernels; otherwise the streamed-operand kernels (csrc/wide.cu) run it."""
    return n % 32 == 0 and n <= 128 and k <= MAX_IN


def supported(linears: Sequence[torch.nn.Linear], x: torch.Tensor) -> bool:
    """True if the stack can run on the tensor-core kernels: CUDA fp32 with biases; every width is covered by one of the
    two kernel families, except a <= 4-wide head behind a hidden layer wider than 128 (tnf_head_bwd's limit; no
    configuration of the reference has one)."""
    if not x.is_cuda or x.dtype != torch.float32 or len(linears) < 2:
        return False
    if any(l.weight.dtype != torch.float32 or l.bias is None for l in linears):
        return False
    if linears[-1].out_features <= 4 and linears[-2].out_features > 128:
        return False
    return True


def _prep(x: torch.Tensor) -> torch.Tensor:
    """2-D, 16-byte aligned rows with a leading dimension multiple of 4 (pads a copy when needed)."""
    x = x.reshape(-1, x.shape[-1])
    if x.stride(1) == 1 and x.stride(0) % 4 == 0 and x.data_ptr() % 16 == 0 and x.stride(0) >= x.size(1):
        return x
    k = x.size(1)
    ld = (k + 3) // 4 * 4
    buf = torch.empty(x.size(0), ld, device=x.device, dtype=torch.float32)
    buf[:, :k].copy_(x)
    return buf[:, :k]


def _lin_fwd(x, w, b, relu, head=None, head_act=0, want_y=True):
    m, k = x.shape
    n = w.size(0)
    if not _resident(n, k):
        y = torch.empty(m, n, device=x.device)
        _lib.call("tnf_wide_linear_fwd", x.data_ptr(), x.stride(0), w.data_ptr(), w.stride(0), _lib.ptr(b), y.data_ptr(), n, m, n, k,
                  1 if relu else 0, _lib.stream_ptr(), nbytes=4 * (m * (k + n) + n * k), flops=2 * m * n * k)
        ho = None
        if head is not None:   # the small output layer as its own launch (the resident kernels fuse it into the epilogue)
            hw, hb = head
            nh = hw.size(0)
            ho = torch.empty(m, nh, device=x.device)
            _lib.call("tnf_wide_linear_fwd", y.data_ptr(), n, hw.data_ptr(), hw.stride(0), hb.data_ptr(), ho.data_ptr(), nh, m, nh, n,
                      _HEAD_ACT_TO_WIDE[head_act], _lib.stream_ptr(), nbytes=4 * m * (n + nh), flops=2 * m * nh * n)
        return y, ho
    y = torch.empty(m, n, device=x.device) if want_y else None
    hw = hb = ho = None
    nh = 0
    if head is not None:
        hw, hb = head
        nh = hw.size(0)
        ho = torch.empty(m, nh, device=x.device)
    _lib.call("tnf_linear_fwd", x.data_ptr(), x.stride(0), w.data_ptr(), _lib.ptr(b), _lib.ptr(y), n, m, n, k,
              int(relu), _lib.ptr(hw), _lib.ptr(hb), _lib.ptr(ho), nh, head_act, _lib.stream_ptr(),
              nbytes=4 * (m * (k + (n if want_y else 0) + nh) + n * k), flops=2 * m * n * (k + nh))
    return y, ho


def _wgrad(dy, x, gw, gb, stream):
    m, (n, k) = dy.size(0), gw.shape
    if _resident(n, k):
        _lib.call("tnf_linear_bwd_weight", dy.data_ptr(), dy.stride(0), x.data_ptr(), x.stride(0), gw.data_ptr(), gb.data_ptr(), m, n, k,
                  stream, nbytes=4 * (m * (n + k) + n * k), flops=2 * m * n * k)
    else:
        _lib.call("tnf_wide_linear_bwd_weight", dy.data_ptr(), dy.stride(0), x.data_ptr(), x.stride(0), gw.data_ptr(), k, gb.data_ptr(),
                  m, n, k, stream, nbytes=4 * (m * (n + k) + n * k), flops=2 * m * n * k)


def _dgrad(dy, w, dx, relu_src, stream, k_begin=0):
    """dx[:, :] = dy @ w[:, k_begin:k_begin + dx.size(1)] (masked by relu_src > 0 when given)."""
    m, n = dy.shape
    k = dx.size(1)
    if _resident(n, w.size(1)) and k_begin == 0 and k == w.size(1):
        _lib.call("tnf_linear_bwd_data", dy.data_ptr(), dy.stride(0), w.data_ptr(), dx.data_ptr(), dx.stride(0),
                  _lib.ptr(relu_src), 0 if relu_src is None else relu_src.stride(0), m, n, k, stream,
                  nbytes=4 * (m * (n + (2 if relu_src is not None else 1) * k) + n * k), flops=2 * m * n * k)
    else:
        _lib.call("tnf_wide_linear_bwd_data", dy.data_ptr(), dy.stride(0), w.data_ptr() + 4 * k_begin, w.stride(0), dx.data_ptr(),
                  dx.stride(0), _lib.ptr(relu_src), 0 if relu_src is None else relu_src.stride(0), m, n, k, stream,
                  nbytes=4 * (m * (n + (2 if relu_src is not None else 1) * k) + n * k), flops=2 * m * n * k)


class _FusedMLP(Function):
    @staticmethod
    def forward(ctx, x, head_act, *params):  # params = w0, b0, w1, b1, ..., w_last, b_last
        _lib.load()
        ws, bs = [p.contiguous() for p in params[0::2]], [p.contiguous() for p in params[1::2]]
        lead = x.shape[:-1]
        x2 = _prep(x.detach())
        n_last = ws[-1].size(0)
        head = n_last <= 4
        acts: List[torch.Tensor] = []
        with torch.cuda.device(x.device):
            h = x2
            n_hidden = len(ws) - 1
            out = None
            for i in range(n_hidden):
                last_hidden = i == n_hidden - 1
                if last_hidden and head:
                    h, out = _lin_fwd(h, ws[i], bs[i], True, head=(ws[-1], bs[-1]), head_act=head_act)
                else:
                    h, _ = _lin_fwd(h, ws[i], bs[i], True)
                acts.append(h)
            if not head:
                out, _ = _lin_fwd(h, ws[-1], bs[-1], False)
        ctx.save_for_backward(x2, out, *acts, *ws)
        ctx.n_layers, ctx.head, ctx.head_act = len(ws), head, head_act
        ctx.x_needs_grad = x.requires_grad
        ctx.in_shape = x.shape
        return out.view(*lead, n_last)

    @staticmethod
    def backward(ctx, grad_out):
        L = ctx.n_layers
        saved = ctx.saved_tensors
        x2, out = saved[0], saved[1]
        acts = list(saved[2:2 + L - 1])
        ws = list(saved[2 + L - 1:])
        m = x2.size(0)
        dev = x2.device
        grad_out = grad_out.reshape(m, -1).contiguous().float()
        # one zero-fill for every weight/bias gradient of the stack (the kernels accumulate with atomics);
        # each slice starts on a 16-byte boundary
        sizes = []
        for w in ws:
            sizes += [w.numel(), w.size(0)]
        offs, tot = [], 0
        for n_el in sizes:
            offs.append(tot)
            tot += (n_el + 3) // 4 * 4
        flat = torch.zeros(tot, device=dev)
        gws = [flat[offs[2 * i]:offs[2 * i] + ws[i].numel()].view_as(ws[i]) for i in range(L)]
        gbs = [flat[offs[2 * i + 1]:offs[2 * i + 1] + ws[i].size(0)] for i in range(L)]
        stream = _lib.stream_ptr()
        with torch.cuda.device(dev):
            h_last = acts[-1]
            if ctx.head:
                dh = torch.empty_like(h_last)
                _lib.call("tnf_head_bwd", h_last.data_ptr(), h_last.stride(0), ws[-1].data_ptr(), out.data_ptr(),
                          grad_out.data_ptr(), dh.data_ptr(), gws[-1].data_ptr(), gbs[-1].data_ptr(), m,
                          h_last.size(1), ws[-1].size(0), ctx.head_act, stream,
                          nbytes=4 * m * (2 * h_last.size(1) + 2 * ws[-1].size(0)))
            else:  # wide last layer without activation
                _wgrad(grad_out, h_last, gws[-1], gbs[-1], stream)
                dh = torch.empty_like(h_last)
                _dgrad(grad_out, ws[-1], dh, h_last, stream)
            gx = None
            for i in range(L - 2, -1, -1):
                inp = acts[i - 1] if i > 0 else x2
                n, k = ws[i].shape
                _wgrad(dh, inp, gws[i], gbs[i], stream)
                if i > 0:
                    dprev = torch.empty_like(inp)
                    _dgrad(dh, ws[i], dprev, inp, stream)
                    dh = dprev
                elif ctx.x_needs_grad:
                    ld = (k + 3) // 4 * 4
                    buf = torch.empty(m, ld, device=dev)
                    _dgrad(dh, ws[i], buf[:, :k], None, stream)
                    gx = buf[:, :k].reshape(ctx.in_shape)
        grads = []
        for gw, gb in zip(gws, gbs):
            grads += [gw, gb]
        return (gx, None, *grads)


def fused_mlp(x: torch.Tensor, linears: Sequence[torch.nn.Linear], head_act: int = 0) -> torch.Tensor:
    params = []
    for l in linears:
        params += [l.weight, l.bias]
    return _FusedMLP.apply(x, head_act, *params)


class _ColorInput(Function):
    """[PE(d) | d | features] rows of the colour head (src/models.py:87) in one pass; gradient flows to features."""

    @staticmethod
    def forward(ctx, features, dirs, n_freqs):
        _lib.load()
        f = features.detach().reshape(-1, features.shape[-1])
        d = dirs.detach().reshape(-1, 3)
        if f.stride(1) != 1:
            f = f.contiguous()
        if d.stride(1) != 1:
            d = d.contiguous()
        n, fd = f.shape
        width = 6 * n_freqs + 3 + fd
        ld = (width + 3) // 4 * 4
        out = torch.empty(n, ld, device=f.device)
        with torch.cuda.device(f.device):
            _lib.call("tnf_color_input", d.data_ptr(), d.stride(0), f.data_ptr(), f.stride(0), n_freqs, fd, out.data_ptr(), ld, n,
                      _lib.stream_ptr(), nbytes=n * (12 + 4 * fd + 4 * ld))
        ctx.fd, ctx.width, ctx.shape = fd, width, features.shape
        return out[:, :width]

    @staticmethod
    def backward(ctx, g):
        return g[:, ctx.width - ctx.fd:ctx.width].reshape(ctx.shape), None, None


def color_input(features: torch.Tensor, dirs: torch.Tensor, n_freqs: int) -> torch.Tensor:
    return _ColorInput.apply(features, dirs, n_freqs)
