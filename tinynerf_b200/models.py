"""Host-side mirror of the reference's models (src/models.py) over the sm_100a kernels.

Same module names, constructor arguments, `feature_dim` attributes and state_dict keys/shapes:

    MLP, PositionalEncoding, truncated_exp                      src/models.py:7-55
    VanillaFeatureMLP / VanillaOpacityDecoder / VanillaColorDecoder   :59-89
    KPlanesFeaturePlane / KPlanesFeatureField                   :93-181
    SawtoothEncoding / CobafaGrid / CobafaFeatureField          :209-266

Feature grids keep their logical [1,C,...] parameter shape but are stored channels-last, which is
what the fused gather kernels (tnf_kplanes_*, tnf_cobafa_*) read: one interpolation corner = one
contiguous C*4-byte line.  The MLPs keep the reference's nn.Linear module tree (state_dict keys) but are
evaluated by the tcgen05 kernels (mlp_ops); there is no cuBLAS / grid_sample / eager path: a shape the kernels do
not cover, or a non-CUDA tensor, raises.
"""
from __future__ import annotations

import ctypes as C
import itertools
from typing import Any, Callable, List, Tuple, cast

import torch
from torch.autograd import Function

from . import _lib, mlp_ops


class MLP(torch.nn.Module):
    """Linear/activation stack with the reference's module tree, hence the same state_dict keys
    (`net.0`, `net.{2+i}.0`, `net.{2+h}`; src/models.py:17-26)."""

    def __init__(self, in_features: int, hidden_features: int, hidden_layers: int,
                 out_features: int | None = None, activation: Callable = torch.nn.ReLU):
        super().__init__()
        Lin, Seq = torch.nn.Linear, torch.nn.Sequential
        blocks: List[torch.nn.Module] = [Lin(in_features, hidden_features), activation()]
        for _ in range(hidden_layers):
            blocks.append(Seq(Lin(hidden_features, hidden_features), activation()))
        blocks.append(Lin(hidden_features, hidden_features if out_features is None else out_features))
        self.net = Seq(*blocks)
        self._relu = activation is torch.nn.ReLU

    def linears(self) -> List[torch.nn.Linear]:
        """The Linear layers in evaluation order (used by the fused MLP kernels)."""
        return [m for m in self.net.modules() if isinstance(m, torch.nn.Linear)]

    def require_supported(self, x: torch.Tensor) -> None:
        """The stack runs on the tcgen05 kernels or not at all (no cuBLAS dispatch): raise for anything they do not cover."""
        _lib.require_cuda(x, "MLP input")
        if not self._relu:
            raise NotImplementedError("MLP: only ReLU activations are implemented on the tensor-core kernels")
        if not mlp_ops.supported(self.linears(), x):
            dims = [self.linears()[0].in_features] + [l.out_features for l in self.linears()]
            raise NotImplementedError(
                f"MLP {dims}: the tcgen05 kernels need fp32 layers with biases, and a <= 4-wide output layer must follow a hidden "
                "layer of at most 128 units; there is no cuBLAS fallback")

    def forward(self, x: torch.Tensor, head_act: int = 0):
        """head_act != 0 asks for the decoder's output activation fused into the last layer (see mlp_ops)."""
        self.require_supported(x)
        return mlp_ops.fused_mlp(x, self.linears(), head_act)


class PositionalEncoding(torch.nn.Module):
    def __init__(self, n_freqs: int):
        super().__init__()
        self.freqs: torch.Tensor
        self.register_buffer("freqs", 2 ** torch.arange(0, n_freqs) * torch.pi)

    def forward(self, x: torch.Tensor) -> torch.Tensor:
        """[..., d] -> [..., d * 2 * n_freqs] (src/models.py:36-39); tnf_positional_encoding.  The frequencies are the
        buffer's 2^k * pi (the kernel rebuilds them as fl32(pi) * 2^k, the same fp32 values)."""
        _lib.load()
        _lib.require_cuda(x, "x")
        d, nf = x.shape[-1], int(self.freqs.numel())
        flat = x.reshape(-1, d).to(torch.float32)
        if flat.stride(-1) != 1:
            flat = flat.contiguous()
        out = torch.empty(flat.size(0), 2 * d * nf, device=x.device)
        with torch.cuda.device(x.device):
            _lib.call("tnf_positional_encoding", flat.data_ptr(), flat.stride(0) if flat.size(0) > 1 else d, d, nf, flat.size(0),
                      out.data_ptr(), out.stride(0), _lib.stream_ptr(), nbytes=4 * flat.size(0) * d * (1 + 2 * nf))
        return out.view(*x.shape[:-1], 2 * d * nf)


class TruncatedExponential(Function):
    """exp forward, exp(clamp(x,-15,15)) backward (src/models.py:42-53)."""

    @staticmethod
    def forward(ctx, x):
        x = x.float()
        ctx.save_for_backward(x)
        return torch.exp(x)

    @staticmethod
    def backward(ctx, g):
        x = ctx.saved_tensors[0]
        return g * torch.exp(torch.clamp(x, min=-15, max=15))


truncated_exp: Callable = TruncatedExponential.apply

"""Vanilla NeRF"""


class VanillaFeatureMLP(torch.nn.Module):
    def __init__(self, n_freqs: int, hidden_features: int, hidden_layers: int):
        super().__init__()
        in_features = n_freqs * 2 * 3
        self.encoding = PositionalEncoding(n_freqs=n_freqs)
        self.net = MLP(in_features, hidden_features, hidden_layers)
        self.feature_dim = hidden_features

    def forward(self, x):
        return self.net(self.encoding(x))


class VanillaOpacityDecoder(torch.nn.Module):
    def __init__(self, feature_dim):
        super().__init__()
        self.net = MLP(feature_dim, 64, 0, 1)
        self.activation = lambda x: truncated_exp(x - 1.0)

    def forward(self, features: torch.Tensor) -> torch.Tensor:
        return self.net(features, head_act=1)  # truncated_exp(x - 1.) fused into the layer epilogue


class VanillaColorDecoder(torch.nn.Module):
    def __init__(self, n_freqs: int, in_features: int, hidden_features: int, hidden_layers: int):
        super().__init__()
        self.pe = PositionalEncoding(n_freqs)
        total_features = in_features + n_freqs * 2 * 3 + 3
        self.net = MLP(total_features, hidden_features, hidden_layers, 3)
        self.activation = torch.nn.Sigmoid()

    def forward(self, features: torch.Tensor, rays_d: torch.Tensor) -> torch.Tensor:
        _lib.require_cuda(features, "features")
        lead = features.shape[:-1]
        f2 = features.reshape(-1, features.shape[-1]).float()
        d2 = rays_d.reshape(-1, 3).float()
        x = mlp_ops.color_input(f2, d2, self.pe.freqs.numel())  # [PE(d) | d | features] in one kernel (src/models.py:87)
        return self.net(x, head_act=2).view(*lead, 3)  # sigmoid fused into the layer epilogue


"""K-Planes https://arxiv.org/abs/2301.10241"""


def _channels_last_param(shape: Tuple[int, ...], init: Callable) -> torch.nn.Parameter:
    """Parameter of logical shape [1,C,*spatial] whose memory is channels-last.  `init` runs on a
    contiguous tensor first so a seeded init gives the same logical values as the reference."""
    tmp = torch.empty(*shape)
    init(tmp)
    perm = (0, *range(2, len(shape)), 1)              # N, spatial..., C
    inv = (0, len(shape) - 1, *range(1, len(shape) - 1))
    storage = tmp.permute(*perm).contiguous()         # physically [1, *spatial, C]
    return torch.nn.Parameter(storage.permute(*inv))  # logical view [1, C, *spatial]


def _channels_last_storage(t: torch.Tensor) -> torch.Tensor:
    """[1,C,*spatial] tensor -> contiguous [*spatial, C] view of the same memory (no copy when the
    tensor already is channels-last, e.g. a parameter created by _channels_last_param)."""
    nd = t.dim()
    v = t.permute(0, *range(2, nd), 1)[0]
    return v if v.is_contiguous() else None  # type: ignore


def _ensure_channels_last_(p: torch.nn.Parameter) -> torch.Tensor:
    v = _channels_last_storage(p.data)
    if v is None:  # e.g. after load_state_dict(assign=True) or .to(memory_format=contiguous)
        nd = p.dim()
        perm = (0, *range(2, nd), 1)
        inv = (0, nd - 1, *range(1, nd - 1))
        p.data = p.data.permute(*perm).contiguous().permute(*inv)
        v = _channels_last_storage(p.data)
    return v


class KPlanesFeaturePlane(torch.nn.Module):
    def __init__(self, feature_dim: int = 8, resolution: Tuple[int, int] = (128, 128),
                 init: Callable = torch.nn.init.uniform_):
        super().__init__()
        self.feature_dim = feature_dim
        self.plane = _channels_last_param((1, feature_dim, *resolution), init)

    def forward(self, x: torch.Tensor) -> torch.Tensor:
        """x: (..., 2).  Stand-alone single-plane lookup (src/models.py:105-113) via tnf_plane_lookup_fwd/bwd; the hot
        path is KPlanesFeatureField.forward, which fuses all nine planes and never calls this."""
        _ensure_channels_last_(self.plane)
        return _GridLookup.apply(x, self.plane)

    def loss_tv(self) -> torch.Tensor:
        """mse(p[1:]-p[:-1]) along H + along W (src/models.py:115-118); one fused kernel (square planes, C % 4 == 0)."""
        _lib.require_cuda(self.plane, "plane")
        if self.plane.shape[2] != self.plane.shape[3] or self.feature_dim % 4 != 0:
            raise NotImplementedError("loss_tv: the TV kernel covers square planes with a multiple of 4 channels")
        _ensure_channels_last_(self.plane)
        return _KPlanesTV.apply(self.feature_dim, self.plane)

    def loss_l1(self) -> torch.Tensor:
        """mean |plane| (src/models.py:120-121); tnf_abs_mean_fwd/bwd."""
        return _AbsMean.apply(self.plane)


class _AbsMean(Function):
    """mean |t| of a dense tensor of any stride order."""

    @staticmethod
    def forward(ctx: Any, t: torch.Tensor):  # type: ignore
        _lib.load()
        _lib.require_cuda(t, "tensor")
        flat = torch.as_strided(t.detach(), (t.numel(),), (1,), t.storage_offset())   # dense memory in storage order
        s = torch.empty(1, dtype=torch.float64, device=t.device)
        with torch.cuda.device(t.device):
            _lib.call("tnf_abs_mean_fwd", flat.data_ptr(), flat.numel(), s.data_ptr(), _lib.stream_ptr(), nbytes=4 * flat.numel())
        ctx.save_for_backward(t)
        return (s / t.numel()).float().reshape(())

    @staticmethod
    def backward(ctx: Any, g: torch.Tensor):  # type: ignore
        (t,) = ctx.saved_tensors
        grad = torch.empty_like(t, memory_format=torch.preserve_format)
        gs = g.detach().float().reshape(1).contiguous()
        with torch.cuda.device(t.device):
            _lib.call("tnf_abs_mean_bwd", t.data_ptr(), t.numel(), gs.data_ptr(), grad.data_ptr(), _lib.stream_ptr(),
                      nbytes=8 * t.numel())
        return grad


class _GridLookup(Function):
    """grid_sample(param[1,C,*spatial], x, bilinear, zeros, align_corners=True) for one channels-last 2-D plane
    (x [...,2]) or 3-D grid (x [...,3]) -> [..., C]; tnf_plane_lookup_* / tnf_grid3_lookup_*."""

    @staticmethod
    def forward(ctx: Any, x: torch.Tensor, param: torch.Tensor):  # type: ignore
        _lib.load()
        _lib.require_cuda(x, "x")
        _lib.require_cuda(param, "grid")
        stor = _channels_last_storage(param)
        if stor is None or param.dtype != torch.float32:
            raise RuntimeError("grid parameters must be fp32 and channels-last")
        nd = param.dim() - 2
        if x.shape[-1] != nd:
            raise RuntimeError(f"expected coordinates with last dimension {nd}")
        x2 = x.detach().reshape(-1, nd).float().contiguous()
        n, ch = x2.size(0), int(param.shape[1])
        out = torch.empty(n, ch, device=x.device)
        with torch.cuda.device(x.device):
            if nd == 2:
                _lib.call("tnf_plane_lookup_fwd", stor.data_ptr(), int(param.shape[2]), int(param.shape[3]), ch, x2.data_ptr(), 2, n,
                          out.data_ptr(), _lib.stream_ptr(), nbytes=n * (8 + 4 * ch))
            else:
                _lib.call("tnf_grid3_lookup_fwd", stor.data_ptr(), int(param.shape[2]), int(param.shape[3]), int(param.shape[4]), ch,
                          x2.data_ptr(), 3, n, out.data_ptr(), _lib.stream_ptr(), nbytes=n * (12 + 4 * ch))
        ctx.save_for_backward(x2, param)
        return out.view(*x.shape[:-1], ch)

    @staticmethod
    def backward(ctx: Any, grad_out: torch.Tensor):  # type: ignore
        x2, param = ctx.saved_tensors
        nd = param.dim() - 2
        n, ch = x2.size(0), int(param.shape[1])
        g = grad_out.reshape(n, ch).float().contiguous()
        grad = torch.zeros_like(param, memory_format=torch.preserve_format)
        gstor = _channels_last_storage(grad)
        with torch.cuda.device(x2.device):
            if nd == 2:
                _lib.call("tnf_plane_lookup_bwd", gstor.data_ptr(), int(param.shape[2]), int(param.shape[3]), ch, x2.data_ptr(), 2, n,
                          g.data_ptr(), _lib.stream_ptr(), nbytes=n * (8 + 4 * ch))
            else:
                _lib.call("tnf_grid3_lookup_bwd", gstor.data_ptr(), int(param.shape[2]), int(param.shape[3]), int(param.shape[4]), ch,
                          x2.data_ptr(), 3, n, g.data_ptr(), _lib.stream_ptr(), nbytes=n * (12 + 4 * ch))
        return None, grad


class _KPlanesTV(Function):
    """mean_i loss_tv(plane_i) for square channels-last planes, via tnf_tv_fwd / tnf_tv_bwd."""

    @staticmethod
    def forward(ctx: Any, channels: int, *planes: torch.Tensor):  # type: ignore
        _lib.load()
        n = len(planes)
        stor = [_channels_last_storage(p) for p in planes]
        if any(v is None for v in stor):
            raise RuntimeError("K-Planes parameters must be channels-last")
        res = [int(p.shape[-1]) for p in planes]
        dev = planes[0].device
        sums = torch.empty(2 * n, dtype=torch.float64, device=dev)
        ptrs = (C.c_void_p * n)(*[t.data_ptr() for t in stor])
        res_arr = (C.c_int32 * n)(*res)
        with torch.cuda.device(dev):
            _lib.call("tnf_tv_fwd", ptrs, res_arr, n, channels, sums.data_ptr(), _lib.stream_ptr(),
                      nbytes=sum(t.numel() for t in stor) * 4)
        denom = torch.tensor([float(channels * (r - 1) * r) for r in res for _ in range(2)], dtype=torch.float64)
        ctx.save_for_backward(*planes)
        ctx.channels, ctx.res = channels, res
        return ((sums / denom.to(dev)).sum() / n).float()

    @staticmethod
    def backward(ctx: Any, grad_out: torch.Tensor):  # type: ignore
        planes = ctx.saved_tensors
        n = len(planes)
        grads = [torch.empty_like(p) for p in planes]  # every element is written by the kernel
        ptrs = (C.c_void_p * n)(*[_channels_last_storage(p).data_ptr() for p in planes])
        gptrs = (C.c_void_p * n)(*[_channels_last_storage(g).data_ptr() for g in grads])
        res_arr = (C.c_int32 * n)(*ctx.res)
        wts = (C.c_float * n)(*[1.0 / n] * n)
        gs = grad_out.detach().float().reshape(1).contiguous()
        with torch.cuda.device(gs.device):
            _lib.call("tnf_tv_bwd", ptrs, gptrs, res_arr, n, ctx.channels, wts, gs.data_ptr(), 0, _lib.stream_ptr(),
                      nbytes=2 * sum(p.numel() for p in planes) * 4)
        return (None, *grads)


class _KPlanesLookup(Function):
    """features[N, S*C] = cat_s prod_p bilinear(plane[s][p], x[(i_p, j_p)])  via tnf_kplanes_fwd/bwd."""

    @staticmethod
    def forward(ctx: Any, x: torch.Tensor, channels: int, *planes: torch.Tensor):  # type: ignore
        _lib.load()
        _lib.require_cuda(x, "x")
        n_scales = len(planes) // 3
        stor = []
        for p in planes:
            _lib.require_cuda(p, "plane")
            v = _channels_last_storage(p)
            if v is None or p.dtype != torch.float32:
                raise RuntimeError("K-Planes parameters must be fp32 and channels-last")
            stor.append(v)
        res = [int(planes[3 * s].shape[-1]) for s in range(n_scales)]
        for s in range(n_scales):
            for p in planes[3 * s:3 * s + 3]:
                if tuple(p.shape) != (1, channels, res[s], res[s]):
                    raise RuntimeError("the fused K-Planes lookup needs square planes of one resolution per scale")
        x2 = x.detach()
        if x2.dim() != 2 or x2.size(1) != 3 or x2.dtype != torch.float32 or x2.stride(1) != 1:
            x2 = x2.reshape(-1, 3).float().contiguous()
        n = x2.size(0)
        out = torch.empty(n, n_scales * channels, device=x.device)
        ptrs = (C.c_void_p * len(stor))(*[t.data_ptr() for t in stor])
        res_arr = (C.c_int32 * n_scales)(*res)
        with torch.cuda.device(x.device):
            pbytes = sum(t.numel() for t in stor) * 4
            _lib.call("tnf_kplanes_fwd", ptrs, res_arr, n_scales, channels, x2.data_ptr(),
                      x2.stride(0) if n > 0 else 3, n, out.data_ptr(), _lib.stream_ptr(),
                      nbytes=n * (12 + 4 * n_scales * channels) + pbytes)
        ctx.save_for_backward(x2, *planes)
        ctx.channels = channels
        ctx.res = res
        ctx.lead_shape = x.shape[:-1]
        return out.view(*x.shape[:-1], n_scales * channels)

    @staticmethod
    def backward(ctx: Any, grad_out: torch.Tensor):  # type: ignore
        _lib.load()
        x2, *planes = ctx.saved_tensors
        n_scales = len(planes) // 3
        channels = ctx.channels
        n = x2.size(0)
        grad_out = grad_out.reshape(n, n_scales * channels).contiguous()
        # one zero-fill for all plane gradients; each gradient is a channels-last view like its parameter
        flat = torch.zeros(sum(p.numel() for p in planes), device=x2.device)
        grads, off = [], 0
        for p in planes:
            grads.append(torch.as_strided(flat, p.shape, p.stride(), off))
            off += p.numel()
        gstor = [_channels_last_storage(g) for g in grads]
        stor = [_channels_last_storage(p) for p in planes]
        ptrs = (C.c_void_p * len(stor))(*[t.data_ptr() for t in stor])
        gptrs = (C.c_void_p * len(gstor))(*[t.data_ptr() for t in gstor])
        res_arr = (C.c_int32 * n_scales)(*ctx.res)
        with torch.cuda.device(x2.device):
            pbytes = sum(t.numel() for t in stor) * 4
            _lib.call("tnf_kplanes_bwd", ptrs, gptrs, res_arr, n_scales, channels, x2.data_ptr(),
                      x2.stride(0) if n > 0 else 3, n, grad_out.data_ptr(), _lib.stream_ptr(),
                      nbytes=n * (12 + 4 * n_scales * channels) + 2 * pbytes)
        return (None, None, *grads)


class KPlanesFeatureField(torch.nn.Module):
    def __init__(self, feature_dim: int = 32):
        super().__init__()
        self.planes = torch.nn.ModuleList([
            torch.nn.ModuleList([KPlanesFeaturePlane(feature_dim, resolution=(r, r)) for _ in range(3)])
            for r in (128, 256, 512)
        ])
        self.dropout = torch.nn.Dropout(0.0)
        # pairs of coordinates used by the three planes of a scale, *in that order* (src/models.py:145)
        self.dimension_pairs = list(itertools.combinations(range(3), 2))
        self.plane_channels = feature_dim
        self.feature_dim = 32 * len(self.planes)  # the reference hard-codes 32 here (src/models.py:146)
        for plane_scale in self.planes:
            assert isinstance(plane_scale, torch.nn.ModuleList)
            assert len(plane_scale) == len(self.dimension_pairs)

    def _plane_params(self) -> List[torch.nn.Parameter]:
        out = []
        for plane_scale in self.planes:
            for plane in plane_scale:  # type: ignore
                _ensure_channels_last_(plane.plane)
                out.append(plane.plane)
        return out

    def forward(self, x: torch.Tensor) -> torch.Tensor:
        """x: (..., 3) in [-1,1] -> (..., 3*C) (src/models.py:153-163); one fused kernel."""
        return self.dropout(_KPlanesLookup.apply(x, self.plane_channels, *self._plane_params()))

    def loss_tv(self) -> torch.Tensor:
        """mean over the nine planes of plane.loss_tv() (src/models.py:165-172); one kernel each way."""
        params = self._plane_params()
        _lib.require_cuda(params[0], "planes")
        if self.plane_channels % 4 != 0:
            raise NotImplementedError("loss_tv: the TV kernel needs a multiple of 4 channels")
        return _KPlanesTV.apply(self.plane_channels, *params)

    def loss_l1(self) -> torch.Tensor:
        loss = 0.0
        count = 0
        for plane_scale in self.planes:
            for plane in plane_scale:  # type: ignore
                loss += plane.loss_l1()
                count += 1
        return cast(torch.Tensor, loss) / count

    @torch.no_grad()
    def add_tv_grad_(self, scale: float) -> None:
        """p.grad += scale * d loss_tv / d p for the nine planes, straight into the existing gradient buffers
        (tnf_tv_bwd with accumulate=1): what autograd would add for a `scale * loss_tv()` term of the loss, without
        nine temporaries and nine accumulation passes."""
        params = self._plane_params()
        n = len(params)
        for p in params:
            if p.grad is None:
                p.grad = torch.zeros_like(p)
            if p.grad.stride() != p.stride():
                p.grad = torch.empty_like(p).copy_(p.grad)
        ptrs = (C.c_void_p * n)(*[_channels_last_storage(p).data_ptr() for p in params])
        gptrs = (C.c_void_p * n)(*[_channels_last_storage(p.grad).data_ptr() for p in params])
        res_arr = (C.c_int32 * n)(*[int(p.shape[-1]) for p in params])
        wts = (C.c_float * n)(*[1.0 / n] * n)
        gs = torch.full((1,), float(scale), device=params[0].device)
        with torch.cuda.device(gs.device):
            _lib.call("tnf_tv_bwd", ptrs, gptrs, res_arr, n, self.plane_channels, wts, gs.data_ptr(), 1, _lib.stream_ptr(),
                      nbytes=3 * sum(p.numel() for p in params) * 4)


"""CoBaFa https://arxiv.org/abs/2302.01226"""


class SawtoothEncoding(torch.nn.Module):
    def __init__(self, f):
        super().__init__()
        self.f = f

    def forward(self, x: torch.Tensor) -> torch.Tensor:
        """2 * ((f x) mod 1) - 1 (src/models.py:213-214).  The fused Cobafa lookup (tnf_cobafa_fwd) applies this inside the
        kernel; as a stand-alone module it is a parameter-free elementwise map kept in torch (three pointwise ops)."""
        return 2.0 * ((self.f * x) % 1.0) - 1.0  # also normalize to [-1, 1]


class CobafaGrid(torch.nn.Module):
    def __init__(self, res: int | Tuple[int, int, int], feature_dim: int, init: Callable = torch.nn.init.uniform_):
        super().__init__()
        resolution = (res, res, res) if isinstance(res, int) else res
        self.grid = _channels_last_param((1, feature_dim, *resolution), init)
        self.feature_dim = feature_dim

    def forward(self, x: torch.Tensor) -> torch.Tensor:
        """x: (..., 3).  Stand-alone lookup (src/models.py:228-238) via tnf_grid3_lookup_fwd/bwd; the hot path is the
        fused CobafaFeatureField.forward."""
        _ensure_channels_last_(self.grid)
        return _GridLookup.apply(x, self.grid)


class _CobafaLookup(Function):
    """features[N, sum c_l] = cat_l trilinear(basis_l, saw_l(x)) * trilinear(coef, x)[l]."""

    @staticmethod
    def _tables(basis, coef, freqs):
        L = len(basis)
        res, ch = [], []
        for b in basis:
            if b.shape[2] != b.shape[3] or b.shape[3] != b.shape[4]:
                raise RuntimeError("the fused Cobafa lookup needs cubic basis grids")
            res.append(int(b.shape[-1]))
            ch.append(int(b.shape[1]))
        if coef.shape[1] != L or not (coef.shape[2] == coef.shape[3] == coef.shape[4]):
            raise RuntimeError("coef grid must be cubic with one channel per level")
        return ((C.c_int32 * L)(*res), (C.c_int32 * L)(*ch), (C.c_float * L)(*[float(f) for f in freqs]), sum(ch))

    @staticmethod
    def forward(ctx: Any, x: torch.Tensor, freqs, coef: torch.Tensor, *basis: torch.Tensor):  # type: ignore
        _lib.load()
        _lib.require_cuda(x, "x")
        L = len(basis)
        res_arr, ch_arr, f_arr, feat = _CobafaLookup._tables(basis, coef, freqs)
        stor = [_channels_last_storage(b) for b in basis]
        cst = _channels_last_storage(coef)
        if cst is None or any(s is None for s in stor):
            raise RuntimeError("Cobafa parameters must be channels-last")
        x2 = x.detach()
        if x2.dim() != 2 or x2.size(1) != 3 or x2.dtype != torch.float32 or x2.stride(1) != 1:
            x2 = x2.reshape(-1, 3).float().contiguous()
        n = x2.size(0)
        out = torch.empty(n, feat, device=x.device)
        ptrs = (C.c_void_p * L)(*[t.data_ptr() for t in stor])
        with torch.cuda.device(x.device):
            pbytes = (sum(t.numel() for t in stor) + cst.numel()) * 4
            _lib.call("tnf_cobafa_fwd", ptrs, res_arr, ch_arr, f_arr, L, cst.data_ptr(), int(coef.shape[-1]),
                      x2.data_ptr(), x2.stride(0) if n > 0 else 3, n, out.data_ptr(), _lib.stream_ptr(),
                      nbytes=n * (12 + 4 * feat) + pbytes)
        ctx.save_for_backward(x2, coef, *basis)
        ctx.freqs = list(freqs)
        return out.view(*x.shape[:-1], feat)

    @staticmethod
    def backward(ctx: Any, grad_out: torch.Tensor):  # type: ignore
        _lib.load()
        x2, coef, *basis = ctx.saved_tensors
        L = len(basis)
        res_arr, ch_arr, f_arr, feat = _CobafaLookup._tables(basis, coef, ctx.freqs)
        n = x2.size(0)
        grad_out = grad_out.reshape(n, feat).contiguous()
        gb = [torch.zeros_like(b) for b in basis]
        gc = torch.zeros_like(coef)
        ptrs = (C.c_void_p * L)(*[_channels_last_storage(b).data_ptr() for b in basis])
        gptrs = (C.c_void_p * L)(*[_channels_last_storage(g).data_ptr() for g in gb])
        with torch.cuda.device(x2.device):
            pbytes = (sum(b.numel() for b in basis) + coef.numel()) * 4
            _lib.call("tnf_cobafa_bwd", ptrs, gptrs, res_arr, ch_arr, f_arr, L,
                      _channels_last_storage(coef).data_ptr(), _channels_last_storage(gc).data_ptr(),
                      int(coef.shape[-1]), x2.data_ptr(), x2.stride(0) if n > 0 else 3, n, grad_out.data_ptr(),
                      _lib.stream_ptr(), nbytes=n * (12 + 4 * feat) + 2 * pbytes)
        return (None, None, gc, *gb)


class CobafaFeatureField(torch.nn.Module):
    def __init__(self, basis_res: List[int | Tuple[int, int, int]], coef_res: int | Tuple[int, int, int],
                 freqs: List[float], channels: List[int], mlp_hidden_dim: int):
        super().__init__()
        assert len(basis_res) == len(freqs) == len(channels)
        L = len(basis_res)
        self.basis_grids = torch.nn.ModuleList([CobafaGrid(res, c) for res, c in zip(basis_res, channels)])
        self.encoders = torch.nn.ModuleList([SawtoothEncoding(f) for f in freqs])
        self.coef_grid = CobafaGrid(coef_res, L)
        self.dropout = torch.nn.Dropout(0.01)
        self.mlp = MLP(sum(channels), mlp_hidden_dim, 5)
        self.feature_dim = mlp_hidden_dim

    def lookup(self, x: torch.Tensor) -> torch.Tensor:
        """The concatenated basis*coef features [.., sum(channels)] (src/models.py:260-264)."""
        for g in [self.coef_grid, *self.basis_grids]:
            _ensure_channels_last_(g.grid)  # type: ignore
        freqs = [enc.f for enc in self.encoders]  # type: ignore
        return _CobafaLookup.apply(x, freqs, self.coef_grid.grid, *[g.grid for g in self.basis_grids])  # type: ignore

    def forward(self, x: torch.Tensor) -> torch.Tensor:
        """x: [..., 3] normalized in [-1, 1] (src/models.py:258-266)."""
        features = self.dropout(self.lookup(x))
        return self.mlp(features)
