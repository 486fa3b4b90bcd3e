"""Drop-in for the reference's pybind11 module `_cuda` (src/cuda.cu:134-137, loaded at src/core.py:7).

Same two callables, argument order, return value and precondition errors; the work is done by the
sm_100a kernels behind `tnf_weights_fwd/bwd` (include/tinynerf_b200.h).  Differences, all additive:
dtype/size checks (the reference trusts them), launch errors are raised instead of ignored
(src/cuda.cu:86), launches go to torch's current stream with a device guard, and `steps` may be any
1-D strided view (the reference needs `.contiguous()`, src/core.py:196).
"""
from __future__ import annotations

import torch

from . import _lib

TRUSTED_PARTITION = 1
NO_EXACT_TERMINATION = 2


def _check_common(sigmas, steps, info, steps_contiguous=False):
    # same order as the reference's CHECK_INPUT sequence (src/cuda.cu:72-74)
    _lib.require_cuda(sigmas, "sigmas")
    _lib.require_contiguous(sigmas, "sigmas")
    _lib.require_cuda(steps, "steps")
    if steps_contiguous:
        _lib.require_contiguous(steps, "steps")
    _lib.require_cuda(info, "info")
    _lib.require_contiguous(info, "info")
    if sigmas.dim() != 1:
        raise RuntimeError("sigmas.dim() == 1")
    if steps.dim() != 1:
        raise RuntimeError("steps.dim() == 1")
    if info.dim() != 2 or info.size(1) != 2:
        raise RuntimeError("info.dim() == 2 && info.size(1) == 2")
    if sigmas.dtype != torch.float32 or steps.dtype != torch.float32:
        raise RuntimeError("expected scalar type Float for sigmas/steps")
    if info.dtype != torch.int32:
        raise RuntimeError("expected scalar type Int for info")
    if steps.size(0) != sigmas.size(0):
        raise RuntimeError("sigmas and steps must have the same length")
    if sigmas.size(0) > 1 and steps.stride(0) < 1:
        raise RuntimeError("steps must have a positive stride")


def weights_fwd(sigmas: torch.Tensor, steps: torch.Tensor, info: torch.Tensor, threshold: float,
                flags: int = 0, _steps_contiguous: bool = False) -> torch.Tensor:
    """tnf_weights_fwd on torch tensors; `steps` may be strided (e.g. packed_samples[:, 6])."""
    _check_common(sigmas, steps, info, _steps_contiguous)
    _lib.load()
    n, r = sigmas.size(0), info.size(0)
    with torch.cuda.device(sigmas.device):
        weights = torch.empty_like(sigmas)
        status = None if flags & TRUSTED_PARTITION else torch.empty(1, dtype=torch.int32, device=sigmas.device)
        _lib.call("tnf_weights_fwd", sigmas.data_ptr(), steps.data_ptr(), steps.stride(0) if n > 1 else 1,
                  info.data_ptr(), float(threshold), weights.data_ptr(), n, r, flags, _lib.ptr(status),
                  _lib.stream_ptr(), nbytes=12 * n + 8 * r, extra_kernels=0 if flags & TRUSTED_PARTITION else 3)
    return weights


def weights_bwd(sigmas: torch.Tensor, steps: torch.Tensor, info: torch.Tensor, weights: torch.Tensor,
                grad_weights: torch.Tensor, flags: int = 0, _steps_contiguous: bool = False) -> torch.Tensor:
    _check_common(sigmas, steps, info, _steps_contiguous)
    for t, name in ((weights, "weights"), (grad_weights, "grad_weights")):
        _lib.require_cuda(t, name)
        _lib.require_contiguous(t, name)
        if t.dim() != 1:
            raise RuntimeError(f"{name}.dim() == 1")
        if t.dtype != torch.float32 or t.size(0) != sigmas.size(0):
            raise RuntimeError(f"{name} must be float32 with the same length as sigmas")
    _lib.load()
    n, r = sigmas.size(0), info.size(0)
    with torch.cuda.device(sigmas.device):
        grad_sigmas = torch.empty_like(sigmas)
        status = None if flags & TRUSTED_PARTITION else torch.empty(1, dtype=torch.int32, device=sigmas.device)
        _lib.call("tnf_weights_bwd", sigmas.data_ptr(), steps.data_ptr(), steps.stride(0) if n > 1 else 1,
                  info.data_ptr(), weights.data_ptr(), grad_weights.data_ptr(), grad_sigmas.data_ptr(), n, r, flags,
                  _lib.ptr(status), _lib.stream_ptr(), nbytes=20 * n + 8 * r,
                  extra_kernels=0 if flags & TRUSTED_PARTITION else 3)
    return grad_sigmas


def compute_weights_fwd(sigmas: torch.Tensor, steps: torch.Tensor, info: torch.Tensor,
                        threshold: float) -> torch.Tensor:
    """Reference signature (src/cuda.cu:66-71): contiguous CUDA inputs, returns fresh [N] weights."""
    return weights_fwd(sigmas, steps, info, threshold, 0, True)


def compute_weights_bwd(sigmas: torch.Tensor, steps: torch.Tensor, info: torch.Tensor, weights: torch.Tensor,
                        grad_weights: torch.Tensor) -> torch.Tensor:
    """Reference signature (src/cuda.cu:97-103)."""
    return weights_bwd(sigmas, steps, info, weights, grad_weights, 0, True)
