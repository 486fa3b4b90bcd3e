"""ctypes wrappers over oracle/tnf_oracle.c (CPU torch tensors in, CPU torch tensors out).
TEST INFRASTRUCTURE ONLY -- see oracle/__init__.py."""
from __future__ import annotations

import ctypes as C

import torch

from . import build

_lib = None


class MarchParams(C.Structure):
    _fields_ = [("scene", C.c_int), ("n_steps", C.c_int), ("aabb", C.c_float * 6), ("near", C.c_float),
                ("far", C.c_float), ("step_size", C.c_float), ("t_table", C.c_void_p),
                ("step_table", C.c_void_p), ("grid", C.c_void_p), ("gd", C.c_int), ("gh", C.c_int),
                ("gw", C.c_int), ("thr", C.c_float), ("noise", C.c_void_p), ("use_fma", C.c_int)]


def lib():
    global _lib
    if _lib is None:
        _lib = C.CDLL(str(build()))
        _lib.orc_march.restype = C.c_int64
        _lib.orc_trilinear.restype = C.c_float
    return _lib


def _f(t):
    assert t.device.type == "cpu" and t.dtype == torch.float32 and t.is_contiguous()
    return C.c_void_p(t.data_ptr())


def _i(t):
    assert t.device.type == "cpu" and t.dtype == torch.int32 and t.is_contiguous()
    return C.c_void_p(t.data_ptr())


def weights_fwd(sigmas, steps, info, threshold):
    sigmas, steps, info = sigmas.float().contiguous(), steps.float().contiguous(), info.int().contiguous()
    out = torch.empty_like(sigmas)
    lib().orc_weights_fwd(_f(sigmas), _f(steps), C.c_int64(1), _i(info), C.c_float(threshold), _f(out),
                          C.c_int64(sigmas.numel()), C.c_int64(info.size(0)))
    return out


def weights_bwd(sigmas, steps, info, weights, grad_weights):
    sigmas, steps, info = sigmas.float().contiguous(), steps.float().contiguous(), info.int().contiguous()
    weights, grad_weights = weights.float().contiguous(), grad_weights.float().contiguous()
    out = torch.empty_like(sigmas)
    lib().orc_weights_bwd(_f(sigmas), _f(steps), C.c_int64(1), _i(info), _f(weights), _f(grad_weights), _f(out),
                          C.c_int64(sigmas.numel()), C.c_int64(info.size(0)))
    return out


def occ_query(grid, coords, thr, use_fma=True):
    grid, coords = grid.float().contiguous(), coords.reshape(-1, 3).float().contiguous()
    n = coords.size(0)
    mask = torch.empty(n, dtype=torch.uint8)
    vals = torch.empty(n)
    D, H, W = grid.shape
    lib().orc_occ_query(_f(grid), D, H, W, _f(coords), C.c_int64(n), C.c_float(thr), int(use_fma),
                        C.c_void_p(mask.data_ptr()), _f(vals))
    return mask.bool(), vals


def march(scene, rays_o, rays_d, n_steps, grid, thr, *, aabb=None, near=0.0, far=1e5, step_size=0.0,
          t_table=None, step_table=None, noise=None, use_fma=True):
    """-> (mask [R,S] bool, info [R,2] int32, packed [N,7])"""
    rays_o, rays_d, grid = rays_o.float().contiguous(), rays_d.float().contiguous(), grid.float().contiguous()
    R = rays_o.size(0)
    P = MarchParams()
    P.scene, P.n_steps = scene, n_steps
    keep = [rays_o, rays_d, grid]
    if aabb is not None:
        for i, v in enumerate(aabb.reshape(-1).tolist()):
            P.aabb[i] = v
    P.near, P.far, P.step_size = near, far, step_size
    if t_table is not None:
        t_table, step_table = t_table.float().contiguous(), step_table.float().contiguous()
        keep += [t_table, step_table]
        P.t_table, P.step_table = t_table.data_ptr(), step_table.data_ptr()
    P.grid = grid.data_ptr()
    P.gd, P.gh, P.gw = grid.shape
    P.thr = thr
    if noise is not None:
        noise = noise.float().contiguous()
        keep.append(noise)
        P.noise = noise.data_ptr()
    P.use_fma = int(use_fma)
    mask = torch.empty(R, n_steps, dtype=torch.uint8)
    info = torch.empty(R, 2, dtype=torch.int32)
    n = lib().orc_march(C.byref(P), _f(rays_o), _f(rays_d), C.c_int64(R), C.c_void_p(mask.data_ptr()), _i(info),
                        None, C.c_int64(0))
    packed = torch.empty(n, 7)
    lib().orc_march(C.byref(P), _f(rays_o), _f(rays_d), C.c_int64(R), None, _i(info), _f(packed), C.c_int64(n))
    return mask.bool(), info, packed


def occ_update_coords(shape, cell0, noise):
    noise = noise.reshape(-1, 3).float().contiguous()
    out = torch.empty_like(noise)
    lib().orc_occ_update_coords(shape[0], shape[1], shape[2], C.c_int64(cell0), C.c_int64(noise.size(0)),
                                _f(noise), _f(out))
    return out


def occ_update_apply(grid, cell0, sigma, step, thr, decay):
    grid = grid.float().contiguous().clone()
    sigma = sigma.reshape(-1).float().contiguous()
    lib().orc_occ_update_apply(_f(grid), C.c_int64(cell0), C.c_int64(sigma.numel()), _f(sigma), C.c_float(step),
                               C.c_float(thr), C.c_float(decay))
    return grid


def composite_fwd(w, rgb, info, bg=None):
    w, rgb, info = w.float().contiguous(), rgb.float().contiguous(), info.int().contiguous()
    out = torch.empty(info.size(0), 3)
    bgp = None if bg is None else (C.c_float * 3)(*[float(v) for v in bg])
    lib().orc_composite_fwd(_f(w), _f(rgb), _i(info), C.c_int64(info.size(0)), bgp, _f(out))
    return out


def composite_bwd(w, rgb, info, go, bg=None):
    w, rgb, info, go = w.float().contiguous(), rgb.float().contiguous(), info.int().contiguous(), go.float().contiguous()
    gw, grgb = torch.zeros_like(w), torch.zeros_like(rgb)
    bgp = None if bg is None else (C.c_float * 3)(*[float(v) for v in bg])
    lib().orc_composite_bwd(_f(w), _f(rgb), _i(info), C.c_int64(info.size(0)), bgp, _f(go), _f(gw), _f(grgb))
    return gw, grgb
