"""Host-side mirror of the reference's rendering core (src/core.py) over the sm_100a kernels.

Same class names, constructor arguments, attributes and return values as the reference, so callers
(src/run.py, the reference tests) work unchanged:

    ContractionMip360 / ContractionAABB      src/core.py:11-33
    RayMarcherUnbounded / RayMarcherAABB     src/core.py:36-90
    OccupancyGrid                            src/core.py:93-156
    RayProvider                              src/core.py:158-188
    NerfWeights                              src/core.py:192-207
    NerfRenderer                             src/core.py:209-267

The per-ray helpers (contractions, marchers) are part of the public surface and run their own small kernels
(tnf_contract, tnf_marcher_aabb: the same device code the fused path uses, csrc/helpers.cu); RayProvider never
calls them on the hot path: it hands their parameters to the fused march/contract/occupancy/pack kernels
(tnf_march_count / tnf_march_pack).  There is no CPU or PyTorch path: CUDA tensors are required and the
C-ABI library must be built; anything the kernels do not cover raises.
"""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass
from functools import cached_property
from typing import Any, Callable, List, Tuple

import torch

from . import _cuda, _lib

# ------------------------------------------------------------------------------------------------
# Scene contraction and ray marching strategies (public helpers, same math as the reference)
# ------------------------------------------------------------------------------------------------


@dataclass
class ContractionMip360:
    order: float | int = float("inf")

    @torch.no_grad()
    def __call__(self, coords: torch.Tensor) -> Tuple[torch.Tensor, None]:
        """Mip-NeRF 360 contraction to [-1,1] (src/core.py:16-20), infinity norm (the reference's default and the only
        order its training loop uses, src/run.py:156); tnf_contract."""
        if self.order != float("inf"):
            raise NotImplementedError("ContractionMip360: only order=inf is implemented (src/run.py:156)")
        return _contract(1, None, coords)[0], None


@dataclass
class ContractionAABB:
    aabb: torch.Tensor  # [2,3]

    @torch.no_grad()
    def __call__(self, coords: torch.Tensor) -> Tuple[torch.Tensor, torch.Tensor]:
        """Affine map of the box to [-1,1] plus inside-box mask (src/core.py:27-31); tnf_contract."""
        return _contract(0, self.aabb, coords)


def _aabb6(aabb: torch.Tensor):
    return (C.c_float * 6)(*aabb.detach().to("cpu", torch.float32).reshape(-1).tolist())


def _contract(scene: int, aabb: torch.Tensor | None, coords: torch.Tensor):
    _lib.load()
    _lib.require_cuda(coords, "coords")
    flat = coords.reshape(-1, 3).to(torch.float32).contiguous()
    out = torch.empty_like(flat)
    mask = torch.empty(flat.size(0), dtype=torch.bool, device=flat.device) if scene == 0 else None
    with torch.cuda.device(flat.device):
        _lib.call("tnf_contract", scene, None if aabb is None else _aabb6(aabb), flat.data_ptr(), flat.size(0),
                  out.data_ptr(), _lib.ptr(mask), _lib.stream_ptr(), nbytes=25 * flat.size(0))
    return out.view(coords.shape), None if mask is None else mask.view(coords.shape[:-1])


Contraction = ContractionMip360 | ContractionAABB


@dataclass
class RayMarcherUnbounded:
    n_samples: int = 200
    near: float = 0.0
    far: float = 1e5
    uniform_range: float = 1.0

    @cached_property
    def step_size(self) -> float:
        return self.uniform_range / self.n_samples

    @torch.no_grad()
    def tables(self, device) -> Tuple[torch.Tensor, torch.Tensor]:
        """Ray-independent t / step tables, computed with the reference's own op sequence
        (src/core.py:52-55) on `device` so the values are bit-identical to what it would produce."""
        f = lambda x: torch.where(x < 0.5, 2 * x, 1 / (2 - 2 * x))
        t_values = torch.linspace(0.0, 1.0 - (1.0 / (self.n_samples + 2)), self.n_samples + 1, device=device)
        t_values = f(t_values) * self.uniform_range + self.near
        step_sizes = t_values[1:] - t_values[:-1]
        return t_values[:-1].contiguous(), step_sizes.contiguous()

    @torch.no_grad()
    def __call__(self, rays_o: torch.Tensor, rays_d: torch.Tensor) -> Tuple[torch.Tensor, torch.Tensor]:
        n_rays = rays_o.size(0)
        t_values, step_sizes = self.tables(rays_o.device)
        t_values = torch.broadcast_to(t_values, (n_rays, self.n_samples))
        step_sizes = torch.broadcast_to(step_sizes, (n_rays, self.n_samples))
        return t_values, step_sizes


@dataclass
class RayMarcherAABB:
    aabb: torch.Tensor
    n_samples: int = 200
    near: float = 0.0
    far: float = 1e5

    @cached_property
    def step_size(self) -> float:
        # a 0-dim tensor on the aabb's device, like the reference (src/core.py:68-70)
        return torch.norm(self.aabb[1] - self.aabb[0]) / self.n_samples

    @torch.no_grad()
    def __call__(self, rays_o: torch.Tensor, rays_d: torch.Tensor) -> Tuple[torch.Tensor, torch.Tensor]:
        """t_values, step_sizes [R, n_samples] (src/core.py:73-88); tnf_marcher_aabb."""
        _lib.load()
        _lib.require_cuda(rays_o, "rays_o")
        _lib.require_cuda(rays_d, "rays_d")
        o, d = rays_o.to(torch.float32).contiguous(), rays_d.to(torch.float32).contiguous()
        R, S = o.size(0), int(self.n_samples)
        t_values = torch.empty(R, S, device=o.device)
        step_sizes = torch.empty(R, S, device=o.device)
        ss = self.step_size
        step = float(ss.item()) if isinstance(ss, torch.Tensor) else _f32(ss)
        with torch.cuda.device(o.device):
            _lib.call("tnf_marcher_aabb", _aabb6(self.aabb), _f32(self.near), _f32(self.far), step, o.data_ptr(), d.data_ptr(),
                      R, S, t_values.data_ptr(), step_sizes.data_ptr(), _lib.stream_ptr(), nbytes=24 * R + 8 * R * S)
        return t_values, step_sizes


RayMarcher = RayMarcherUnbounded | RayMarcherAABB


def tag_partition(info: torch.Tensor) -> None:
    """Mark `info` as a 0-based contiguous partition of the packed samples (what RayProvider produces): the weights
    kernels then skip their in-kernel validation (TNF_W_TRUSTED_PARTITION).  The tag records the tensor's version counter,
    so an in-place edit afterwards (the reference's own loop does `info[:, 0] += current_size`, src/run.py:236) voids it."""
    info._tnf_partition = info._version


def is_trusted_partition(info: torch.Tensor) -> bool:
    tag = getattr(info, "_tnf_partition", None)
    return tag is not None and tag is not False and tag == info._version


def tag_steps(packed: torch.Tensor, steps: torch.Tensor) -> None:
    """Attach the contiguous copy of column 6 the pack kernel wrote beside the rows; honoured only while neither tensor
    has been modified in place since."""
    packed._tnf_steps = steps
    packed._tnf_steps_versions = (packed._version, steps._version)


def tagged_steps(packed: torch.Tensor) -> torch.Tensor | None:
    steps = getattr(packed, "_tnf_steps", None)
    if steps is None or getattr(packed, "_tnf_steps_versions", None) != (packed._version, steps._version):
        return None
    return steps


def _f32(x) -> float:
    """Python float -> the fp32 value torch uses when the scalar meets a float32 tensor (round to nearest even)."""
    return C.c_float(float(x)).value


# ------------------------------------------------------------------------------------------------
# Occupancy grid
# ------------------------------------------------------------------------------------------------


class OccupancyGrid(torch.nn.Module):
    """Float occupancy grid with trilinear lookup and jittered decay update (src/core.py:93-156).

    `jitter_source`: "cpu" draws the update jitter slice by slice from torch's CPU generator exactly
    like the reference (src/core.py:137), "device" uses the kernels' Philox stream (no host work).
    `slices_per_call`: how many depth slices go through one sigma_fn call (reference: 1).
    """

    def __init__(self, size: List[int] | int, step_size: float, threshold: float = 0.01, decay: float = 0.95):
        super().__init__()
        size = size if isinstance(size, List) else [size, size, size]
        self.decay = decay
        self.step_size = step_size
        self.base_threshold = threshold
        self.grid: torch.Tensor
        self.register_buffer("grid", torch.ones(size, dtype=torch.float))
        self.size = torch.tensor(size, dtype=torch.float)
        self._mean: float | None = 1.0           # host value; None while only the device value is current
        self._mean_t: torch.Tensor | None = None  # device scalar written by update() (no host sync)
        self._thr_t: torch.Tensor | None = None   # device scalar min(base_threshold, mean) for the march kernels
        # kept for interface parity (the reference exposes .coords); the kernels derive cell
        # coordinates from the flat index instead of reading this tensor
        self.coords = torch.flip(torch.stack(torch.meshgrid([
            torch.arange(size[0], dtype=torch.float),
            torch.arange(size[1], dtype=torch.float),
            torch.arange(size[2], dtype=torch.float),
        ], indexing="ij"), -1), [-1])
        self.jitter_source = "cpu"
        self.slices_per_call = size[0]
        self._update_calls = 0
        # "certainly empty" classifier bitfields for the march kernel (tnf_occ_build_empty_bits): rebuilt lazily whenever the
        # grid or the threshold changed; two buffers alternate so a kernel in flight on another stream keeps reading a
        # consistent image while the next one is built.  OFF by default: bit-identical masks, but measured slower than the
        # plain kernel on the bench scene (csrc/march.cu)
        self.use_empty_bits = False
        self._bits_gen = 0
        self._bits_key = None
        self._bits_bufs: List[torch.Tensor] = []
        self._bits_cur = 0

    @property
    def mean(self) -> float:
        """grid.mean() after the last update (src/core.py:145).  update() leaves it on the device; reading it here is what
        synchronises, so a training loop that never looks at it never waits for the update."""
        if self._mean is None:
            self._mean = float(self._mean_t.item())  # type: ignore
        return self._mean

    @mean.setter
    def mean(self, value: float) -> None:
        self._mean, self._mean_t, self._thr_t = float(value), None, None
        self._bits_gen += 1

    def invalidate(self) -> None:
        """Call after writing `grid` through anything torch does not version (a raw pointer, a collective)."""
        self._bits_gen += 1

    def empty_bits(self) -> torch.Tensor | None:
        """The classifier bitfield for the current grid and threshold, built on the current stream when stale
        (32 KB coarse + 256 KB fine for 128^3; None when disabled or when the coarse level would not fit the march kernel's
        shared-memory staging)."""
        if not self.use_empty_bits or not self.grid.is_cuda:
            return None
        D, H, W = self.grid.shape
        n_words = int(_lib.load().tnf_occ_empty_bits_words(D, H, W))
        if ((D + 1) // 2) * ((H + 1) // 2) * ((((W + 1) // 2) + 31) // 32) * 4 > 64 * 1024:   # coarse level must fit the staging
            return None
        key = (self.grid.data_ptr(), self.grid._version, self._bits_gen, self._update_calls)
        if key != self._bits_key:
            if len(self._bits_bufs) < 2 or self._bits_bufs[0].device != self.grid.device or self._bits_bufs[0].numel() != n_words:
                self._bits_bufs = [torch.empty(n_words, dtype=torch.int32, device=self.grid.device) for _ in range(2)]
            self._bits_cur ^= 1
            bits = self._bits_bufs[self._bits_cur]
            thr_t = self.threshold_tensor()
            thr = _f32(self.base_threshold) if thr_t is not None else _f32(self.threshold)
            with torch.cuda.device(self.grid.device):
                _lib.call("tnf_occ_build_empty_bits", self.grid.data_ptr(), D, H, W, thr, _lib.ptr(thr_t), bits.data_ptr(),
                          _lib.stream_ptr(), nbytes=4 * self.grid.numel() + 4 * n_words)
            self._bits_key = key
        return self._bits_bufs[self._bits_cur]

    def threshold_tensor(self) -> torch.Tensor | None:
        """Device scalar min(base_threshold, mean) when the mean currently lives on the device, else None."""
        if self._mean_t is None:
            return None
        if self._thr_t is None:
            self._thr_t = torch.clamp(self._mean_t.float(), max=_f32(self.base_threshold)).reshape(1)
        return self._thr_t

    def _set_mean_from_grid(self) -> None:
        self._mean, self._thr_t = None, None
        self._mean_t = self.grid.mean()
        self._bits_gen += 1

    @torch.no_grad()
    def occupancy(self) -> float:
        return (self.grid > self.threshold).sum().item() / self.grid.numel()

    @property
    def threshold(self) -> float:
        return min(self.base_threshold, self.mean)

    @property
    def device(self) -> torch.device:
        return self.grid.device

    def _step_size_f32(self) -> float:
        s = self.step_size
        if isinstance(s, torch.Tensor):  # a 0-dim device tensor in the reference (src/core.py:68-70): read it once
            key = (s.data_ptr(), s._version)
            if getattr(self, "_step_cache", (None, None))[0] != key:
                self._step_cache = (key, float(s.item()))
            return self._step_cache[1]
        return _f32(s)

    @torch.no_grad()
    def update(self, sigma_fn: Callable[[torch.Tensor], torch.Tensor], noise: torch.Tensor | None = None):
        """grid = where(1-exp(-sigma*step) > thr, 1, decay*grid) at jittered cell positions, then
        mean = grid.mean() (src/core.py:134-145).

        noise: optional [D,H,W,3] U[0,1) tensor (any device) replacing the internally drawn jitter.
        """
        self.update_slices(sigma_fn, 0, self.grid.size(0), noise)
        self._set_mean_from_grid()

    @torch.no_grad()
    def update_slices(self, sigma_fn: Callable[[torch.Tensor], torch.Tensor], z_begin: int, z_end: int,
                      noise: torch.Tensor | None = None):
        """The update restricted to depth slices [z_begin, z_end) (a rank's shard); does not touch
        `mean`.  The threshold used is the one in force when the call starts, as in the reference."""
        _lib.load()
        _lib.require_cuda(self.grid, "occupancy grid")
        D, H, W = self.grid.shape
        dev = self.device
        per_slice = H * W
        # the threshold in force when the call starts (src/core.py:127,141): the device scalar left by the previous
        # update when there is one (no host sync), else the host value
        thr_t = self.threshold_tensor()
        thr = _f32(self.base_threshold) if thr_t is not None else _f32(self.threshold)
        decay, step = _f32(self.decay), self._step_size_f32()
        spc = max(1, min(int(self.slices_per_call), D))
        cpu_jitter = noise is None and self.jitter_source == "cpu"
        if cpu_jitter:  # keep the CPU generator stream aligned with the reference for skipped slices
            for _ in range(0, z_begin):
                torch.rand(H, W, 3)
        with torch.cuda.device(dev):
            stream = _lib.stream_ptr()
            for z0 in range(z_begin, z_end, spc):
                z1 = min(z_end, z0 + spc)
                n = (z1 - z0) * per_slice
                if noise is not None:
                    u = noise[z0:z1].reshape(-1, 3).to(dev, torch.float32).contiguous()
                elif cpu_jitter:
                    # one CPU draw per depth slice, same generator consumption as the reference
                    u = torch.stack([torch.rand(H, W, 3) for _ in range(z0, z1)]).view(-1, 3).to(dev)
                else:
                    u = None
                ws = getattr(self, "_coords_ws", None)   # grow-only: the update allocates nothing in steady state
                if ws is None or ws.size(0) < n or ws.device != dev:
                    ws = self._coords_ws = torch.empty(n, 3, device=dev)
                coords = ws[:n]
                _lib.call("tnf_occ_update_coords", D, H, W, z0 * per_slice, n, _lib.ptr(u), 0x7E57,
                          self._update_calls * D * per_slice, coords.data_ptr(), stream, nbytes=12 * n)
                sigma = sigma_fn(coords).reshape(-1).float().contiguous()
                if sigma.numel() != n:
                    raise RuntimeError(f"sigma_fn returned {sigma.numel()} values for {n} coordinates")
                _lib.call("tnf_occ_update_apply_dev", self.grid.data_ptr(), z0 * per_slice, n, sigma.data_ptr(),
                          step, thr, _lib.ptr(thr_t), decay, stream, nbytes=12 * n)
        if cpu_jitter:
            for _ in range(z_end, D):
                torch.rand(H, W, 3)
        self._update_calls += 1

    @torch.no_grad()
    def forward(self, coords: torch.Tensor) -> torch.Tensor:
        """coords: [..., 3] in [-1,1] -> bool [...]: trilinear(grid) > threshold (src/core.py:148-156)."""
        _lib.load()
        _lib.require_cuda(self.grid, "occupancy grid")
        new_shape = coords.shape[:-1]
        flat = coords.reshape(-1, 3).to(self.device, torch.float32).contiguous()
        out = torch.empty(flat.size(0), dtype=torch.bool, device=self.device)
        D, H, W = self.grid.shape
        with torch.cuda.device(self.device):
            _lib.call("tnf_occ_query", self.grid.data_ptr(), D, H, W, flat.data_ptr(), flat.size(0),
                      _f32(self.threshold), out.data_ptr(), None, _lib.stream_ptr(), nbytes=13 * flat.size(0))
        return out.view(new_shape)


# ------------------------------------------------------------------------------------------------
# Ray provider: march + contract + occupancy mask + pack, fused
# ------------------------------------------------------------------------------------------------


@dataclass
class RayProvider:
    occupancy_grid: OccupancyGrid
    contraction: Contraction
    ray_marcher: RayMarcher

    def _params(self, device, n_steps: int) -> _lib.MarchParams:
        """tnf_march_params for this provider.  The scene part (aabb / tables / step size) needs device->host reads of
        0-dim tensors, so it is built once per (device, n_steps, scene tensors) and cached; the occupancy state
        (grid pointer, threshold) is refreshed on every call."""
        grid = self.occupancy_grid.grid
        _lib.require_cuda(grid, "occupancy grid")
        if grid.device != device:
            raise RuntimeError("rays and occupancy grid must live on the same device")
        marcher, contraction = self.ray_marcher, self.contraction
        aabb_t = getattr(contraction, "aabb", None)
        key = (str(device), n_steps, id(marcher), id(contraction), None if aabb_t is None else (aabb_t.data_ptr(), aabb_t._version),
               getattr(marcher, "near", None), getattr(marcher, "far", None), getattr(marcher, "uniform_range", None))
        cache = self.__dict__.setdefault("_param_cache", {})
        base = cache.get(key)
        if base is None:
            base = self._scene_params(device, n_steps)
            cache.clear()
            cache[key] = base
        p = _lib.MarchParams()
        C.memmove(C.byref(p), C.byref(base), C.sizeof(p))
        p._keep = base._keep
        p.grid = grid.data_ptr()
        p.gd, p.gh, p.gw = grid.shape
        thr_t = self.occupancy_grid.threshold_tensor()
        if thr_t is not None:   # the mean of the last update is still on the device: hand the kernels the device scalar
            p.threshold_dev = thr_t.data_ptr()
            p.threshold = _f32(self.occupancy_grid.base_threshold)
            p._keep = list(p._keep) + [thr_t]
        else:
            p.threshold = _f32(self.occupancy_grid.threshold)
        bits = self.occupancy_grid.empty_bits()
        if bits is not None:
            p.empty_bits = bits.data_ptr()
            p._keep = list(p._keep) + [bits]
        return p

    def _scene_params(self, device, n_steps: int) -> _lib.MarchParams:
        p = _lib.MarchParams()
        p.n_steps = n_steps
        keep = []  # tensors that must outlive the kernel launches
        if isinstance(self.ray_marcher, RayMarcherAABB):
            if not isinstance(self.contraction, ContractionAABB):
                raise NotImplementedError("RayMarcherAABB is fused with ContractionAABB only")
            aabb = self.contraction.aabb.detach().to("cpu", torch.float32)
            m_aabb = self.ray_marcher.aabb.detach().to("cpu", torch.float32)
            if not torch.equal(aabb, m_aabb):
                raise NotImplementedError("marcher and contraction must share one aabb")
            p.scene = 0
            for i, v in enumerate(aabb.reshape(-1).tolist()):
                p.aabb[i] = v
            p.near, p.far = _f32(self.ray_marcher.near), _f32(self.ray_marcher.far)
            ss = self.ray_marcher.step_size
            p.step_size = float(ss.item()) if isinstance(ss, torch.Tensor) else _f32(ss)
        elif isinstance(self.ray_marcher, RayMarcherUnbounded):
            if not isinstance(self.contraction, ContractionMip360) or self.contraction.order != float("inf"):
                raise NotImplementedError("RayMarcherUnbounded is fused with ContractionMip360(order=inf) only")
            p.scene = 1
            t_tab, s_tab = self.ray_marcher.tables(device)
            p.t_table, p.step_table = t_tab.data_ptr(), s_tab.data_ptr()
            keep += [t_tab, s_tab]
        else:
            raise NotImplementedError(f"unknown ray marcher {type(self.ray_marcher)}")
        p._keep = keep
        return p

    @torch.no_grad()
    def count(self, rays_o: torch.Tensor, rays_d: torch.Tensor, training: bool, noise: torch.Tensor | None = None,
              info_offset: int = 0):
        """First half of __call__: evaluate the sample lattice, return a handle holding the keep-mask bitfield,
        the packing info [R,2] and the packed-sample total (device int64) -- no host sync."""
        _lib.load()
        _lib.require_cuda(rays_o, "rays_o")
        _lib.require_cuda(rays_d, "rays_d")
        dev = rays_o.device
        rays_o = rays_o.to(torch.float32).contiguous()
        rays_d = rays_d.to(torch.float32).contiguous()
        R, S = rays_o.size(0), int(self.ray_marcher.n_samples)
        with torch.cuda.device(dev):
            p = self._params(dev, S)
            if training:
                if noise is None:
                    noise = torch.rand(R, S, device=dev)
                noise = noise.to(dev, torch.float32).contiguous()
                if noise.numel() != R * S:
                    raise RuntimeError("noise must have n_rays*n_samples elements")
                p.jitter, p.noise = 1, noise.data_ptr()
            words = (S + 31) // 32
            mask_bits = torch.empty(max(R, 1) * words, dtype=torch.int32, device=dev)
            info = torch.empty(R, 2, dtype=torch.int32, device=dev)
            n_dev = torch.empty(1, dtype=torch.int64, device=dev)
            _lib.call("tnf_march_count", C.byref(p), rays_o.data_ptr(), rays_d.data_ptr(), R, info_offset,
                      mask_bits.data_ptr(), info.data_ptr(), n_dev.data_ptr(), _lib.stream_ptr(),
                      nbytes=24 * R + 8 * R + 4 * R * words)
        return dict(p=p, rays_o=rays_o, rays_d=rays_d, noise=noise, mask_bits=mask_bits, info=info, n_dev=n_dev,
                    info_offset=info_offset, words=words)

    @torch.no_grad()
    def pack(self, h, n_rays: int | None = None, n: int | None = None):
        """Second half: write the packed samples of the first `n_rays` rays of a count() handle (`n` = their
        packed-sample total; read from the device when omitted)."""
        dev = h["rays_o"].device
        R = h["rays_o"].size(0) if n_rays is None else n_rays
        if n is None:
            n = int(h["n_dev"].item())
        info = h["info"][:R]
        with torch.cuda.device(dev):
            # row capacity rounded up to 16 Ki samples: batch sizes wander by a few percent, and a request a little larger
            # than every cached block makes the caching allocator cudaMalloc (which waits for all queued GPU work)
            cap = (n + 16383) & ~16383
            packed = torch.empty(cap, 7, device=dev)[:n]
            steps = torch.empty(cap, device=dev)[:n]
            _lib.call("tnf_march_pack", C.byref(h["p"]), h["rays_o"].data_ptr(), h["rays_d"].data_ptr(), R,
                      h["info_offset"], h["mask_bits"].data_ptr(), info.data_ptr(), packed.data_ptr(), steps.data_ptr(),
                      None, n, _lib.stream_ptr(), nbytes=24 * R + 8 * R + 4 * R * h["words"] + 32 * n)
        tag_steps(packed, steps)
        if h["info_offset"] == 0:
            tag_partition(info)
        return packed, info

    @torch.no_grad()
    def pack_range(self, h, r0: int, r1: int, sample_offset: int, n: int, packed_buf: torch.Tensor | None = None,
                   steps_buf: torch.Tensor | None = None):
        """Packed samples of the rays [r0, r1) of a count() handle taken WITHOUT jitter (training=False): `sample_offset`
        is the packed index of ray r0's first sample in the handle's numbering, `n` the packed-sample total of the range
        (both known on the host from the handle's info).  Returns (packed [n,7], info [r1-r0,2] with 0-based starts): what
        RayProvider()(rays_o[r0:r1], rays_d[r0:r1], training=False) returns, without re-marching the lattice and without a
        host sync.  `packed_buf` [>=n,7] / `steps_buf` [>=n] are optional fixed-capacity destinations (the render loop of
        src/run.py:34-42 reuses one pair for every chunk of an image)."""
        if h["noise"] is not None:
            raise RuntimeError("pack_range needs a handle counted with training=False (jitter is indexed by ray)")
        dev = h["rays_o"].device
        R, words = r1 - r0, h["words"]
        with torch.cuda.device(dev):
            if packed_buf is None:
                cap = (n + 16383) & ~16383
                packed_buf, steps_buf = torch.empty(cap, 7, device=dev), torch.empty(cap, device=dev)
            if packed_buf.size(0) < n or steps_buf.numel() < n:
                raise RuntimeError("pack_range: destination buffers are smaller than the packed-sample count")
            packed, steps = packed_buf[:n], steps_buf[:n]
            info = h["info"][r0:r1].clone()
            if n > 0:
                _lib.call("tnf_march_pack", C.byref(h["p"]), h["rays_o"][r0:].data_ptr(), h["rays_d"][r0:].data_ptr(), R,
                          h["info_offset"] + sample_offset, h["mask_bits"][r0 * words:].data_ptr(), info.data_ptr(),
                          packed.data_ptr(), steps.data_ptr(), None, n, _lib.stream_ptr(),
                          nbytes=24 * R + 8 * R + 4 * R * words + 32 * n)
            info[:, 0] -= h["info_offset"] + sample_offset
        tag_steps(packed, steps)
        tag_partition(info)
        return packed, info

    @torch.no_grad()
    def __call__(self, rays_o: torch.Tensor, rays_d: torch.Tensor, training: bool,
                 noise: torch.Tensor | None = None, info_offset: int = 0):
        """-> (packed_samples [N,7], packing_info [R,2] int32), same contents and order as the
        reference (src/core.py:165-188).  packed_samples carries two extra attributes used by
        NerfRenderer: `_tnf_steps` (contiguous copy of column 6) and packing_info `_tnf_partition`.

        noise: optional [R,S] jitter in [0,1); default in training is torch.rand on the rays' device,
        which consumes torch's CUDA generator exactly like the reference's rand_like (src/core.py:173).
        One host sync (the packed-sample count); the reference has two.
        """
        return self.pack(self.count(rays_o, rays_d, training, noise, info_offset))


# ------------------------------------------------------------------------------------------------
# Rendering: weights op, compositing, renderer
# ------------------------------------------------------------------------------------------------


class NerfWeights(torch.autograd.Function):
    """w_k = T_k (1 - exp(-sigma_k delta_k)) over packed rays (src/core.py:192-207)."""

    @staticmethod
    def forward(ctx: Any, sigmas: torch.Tensor, steps: torch.Tensor, info: torch.Tensor, threshold: float):  # type: ignore
        sigmas = sigmas.contiguous()
        info = info.contiguous()
        flags = _cuda.TRUSTED_PARTITION if is_trusted_partition(info) else 0
        weights = _cuda.weights_fwd(sigmas, steps, info, threshold, flags)  # steps may be a strided view
        ctx.save_for_backward(sigmas, steps, info, weights)
        ctx.flags = flags
        return weights

    @staticmethod
    def backward(ctx: Any, grad_weights: torch.Tensor):  # type: ignore
        grad_weights = grad_weights.contiguous()
        sigmas, steps, info, weights = ctx.saved_tensors
        grad_sigmas = _cuda.weights_bwd(sigmas, steps, info, weights, grad_weights, ctx.flags)
        return grad_sigmas, None, None, None


class Composite(torch.autograd.Function):
    """rgb_ray = sum_k w_k rgb_k (+ bg (1 - sum_k w_k)) per packed ray (src/core.py:256-265)."""

    @staticmethod
    def forward(ctx: Any, weights: torch.Tensor, rgbs: torch.Tensor, info: torch.Tensor, bg):  # type: ignore
        _lib.load()
        weights, rgbs, info = weights.contiguous(), rgbs.contiguous(), info.contiguous()
        n, r = weights.size(0), info.size(0)
        out = torch.empty(r, 3, device=weights.device)
        bg_arr = None if bg is None else (C.c_float * 3)(*[float(v) for v in bg])
        with torch.cuda.device(weights.device):
            _lib.call("tnf_composite_fwd", weights.data_ptr(), rgbs.data_ptr(), info.data_ptr(), n, r, bg_arr,
                      out.data_ptr(), None, _lib.stream_ptr(), nbytes=16 * n + 20 * r)
        ctx.save_for_backward(weights, rgbs, info)
        ctx.bg = bg
        return out

    @staticmethod
    def backward(ctx: Any, grad_out: torch.Tensor):  # type: ignore
        _lib.load()
        weights, rgbs, info = ctx.saved_tensors
        grad_out = grad_out.contiguous()
        n, r = weights.size(0), info.size(0)
        gw = torch.empty_like(weights) if ctx.needs_input_grad[0] else None
        grgb = torch.empty_like(rgbs) if ctx.needs_input_grad[1] else None
        bg_arr = None if ctx.bg is None else (C.c_float * 3)(*[float(v) for v in ctx.bg])
        if gw is not None or grgb is not None:
            with torch.cuda.device(weights.device):
                _lib.call("tnf_composite_bwd", weights.data_ptr(), rgbs.data_ptr(), info.data_ptr(), n, r,
                          bg_arr, grad_out.data_ptr(), _lib.ptr(gw), _lib.ptr(grgb), _lib.stream_ptr(),
                          nbytes=32 * n + 20 * r)
        return gw, grgb, None, None


class NerfRenderer(torch.nn.Module):
    def __init__(self, feature_module: torch.nn.Module, sigma_decoder: torch.nn.Module,
                 rgb_decoder: torch.nn.Module, bg_color: torch.Tensor | None = None):
        super().__init__()
        self.feature_module = feature_module
        self.sigma_decoder = sigma_decoder
        self.rgb_decoder = rgb_decoder
        self.bg_color = bg_color
        # The colour head is evaluated on EVERY sample instead of the gathered (weights > 0) subset of src/core.py:243-250.
        # Same result (a masked-out sample has weight 0, so it adds exactly 0 to the ray and receives exactly 0
        # gradient) without the nonzero() host sync and the gather/scatter passes; there is no other branch.
        self.dense_rgb = True
        assert hasattr(self.feature_module, "feature_dim"), "feature module requires a feature_dim attribute"

    def forward(self, packed_samples: torch.Tensor, packing_info: torch.Tensor,
                early_termination_threshold: float = 1e-4) -> torch.Tensor:
        """packed [N,7], info [R,2] -> rendered rgb [R,3] (src/core.py:225-267)."""
        device = packed_samples.device
        n_samples = packed_samples.size(0)
        n_rays = packing_info.size(0)
        steps = tagged_steps(packed_samples)
        if steps is None:
            steps = packed_samples[:, 6]  # strided view, read in place by the kernel
        try:
            if n_samples == 0:
                raise ValueError("no samples remaining")
            samples_features = self.feature_module(packed_samples[:, :3])
            from . import heads_ops
            if heads_ops.supported(self.sigma_decoder, self.rgb_decoder, samples_features):
                # both decoders of the reference's shapes in the fused head kernels (one forward kernel, one data-gradient
                # kernel); same values as the two module calls below
                sig, samples_rgbs = heads_ops.fused_heads(self.sigma_decoder, self.rgb_decoder, samples_features,
                                                          packed_samples[:, 3:6])
                samples_sigmas = sig.ravel()
                weights: torch.Tensor = NerfWeights.apply(samples_sigmas, steps, packing_info,
                                                          early_termination_threshold)  # type: ignore
            else:
                samples_sigmas = self.sigma_decoder(samples_features).ravel()
                weights = NerfWeights.apply(samples_sigmas, steps, packing_info, early_termination_threshold)  # type: ignore
                samples_rgbs = self.rgb_decoder(samples_features, packed_samples[:, 3:6])
        except ValueError:
            print("Empty iteration, every sample is masked")
            samples_rgbs = torch.zeros((n_samples, 3), device=device, requires_grad=True)
            weights = torch.zeros(n_samples, device=device, requires_grad=True)
        bg = None if self.bg_color is None else self.bg_color.detach().reshape(-1).tolist()
        if n_rays == 0:
            return torch.zeros((0, 3), device=device)
        return Composite.apply(weights, samples_rgbs, packing_info, bg)
