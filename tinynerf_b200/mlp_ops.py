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

MAX_IN = 160


def supported(linears: Sequence[torch.nn.Linear], x: torch.Tensor) -> bool:
    """True if the stack can run on the tensor-core kernels (CUDA fp32, hidden widths multiple of 32 <= 128,
    inputs <= 160 wide)."""
    if not x.is_cuda or x.dtype != torch.float32 or len(linears) < 2:
        return False
    if any(l.weight.dtype != torch.float32 or l.bias is None for l in linears):
        return False
    hidden = [l.out_features for l in linears[:-1]]
    if any(h % 32 != 0 or h > 128 for h in hidden):
        return False
    if linears[0].in_features > MAX_IN:
        return False
    last = linears[-1].out_features
    return last <= 4 or (last % 32 == 0 and last <= 128)


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
                n, k = ws[-1].shape
                _lib.call("tnf_linear_bwd_weight", grad_out.data_ptr(), grad_out.stride(0), h_last.data_ptr(),
                          h_last.stride(0), gws[-1].data_ptr(), gbs[-1].data_ptr(), m, n, k, stream,
                          nbytes=4 * (m * (n + k) + n * k), flops=2 * m * n * k)
                dh = torch.empty_like(h_last)
                _lib.call("tnf_linear_bwd_data", grad_out.data_ptr(), grad_out.stride(0), ws[-1].data_ptr(), dh.data_ptr(),
                          dh.stride(0), h_last.data_ptr(), h_last.stride(0), m, n, k, stream,
                          nbytes=4 * (m * (n + 2 * k) + n * k), flops=2 * m * n * k)
            gx = None
            for i in range(L - 2, -1, -1):
                inp = acts[i - 1] if i > 0 else x2
                n, k = ws[i].shape
                _lib.call("tnf_linear_bwd_weight", dh.data_ptr(), dh.stride(0), inp.data_ptr(), inp.stride(0),
                          gws[i].data_ptr(), gbs[i].data_ptr(), m, n, k, stream,
                          nbytes=4 * (m * (n + k) + n * k), flops=2 * m * n * k)
                if i > 0:
                    dprev = torch.empty_like(inp)
                    _lib.call("tnf_linear_bwd_data", dh.data_ptr(), dh.stride(0), ws[i].data_ptr(), dprev.data_ptr(),
                              dprev.stride(0), inp.data_ptr(), inp.stride(0), m, n, k, stream,
                              nbytes=4 * (m * (n + 2 * k) + n * k), flops=2 * m * n * k)
                    dh = dprev
                elif ctx.x_needs_grad:
                    ld = (k + 3) // 4 * 4
                    buf = torch.empty(m, ld, device=dev)
                    _lib.call("tnf_linear_bwd_data", dh.data_ptr(), dh.stride(0), ws[i].data_ptr(), buf.data_ptr(), ld,
                              None, 0, m, n, k, stream, nbytes=4 * (m * (n + k) + n * k), flops=2 * m * n * k)
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
