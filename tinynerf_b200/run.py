"""Thin host-side caller of the hot path: the training iteration of the reference's loop
(src/run.py:213-261) and its chunked render (src/run.py:15-50), one process per GPU.

What is kept from the reference: model/marcher/grid/optimiser construction and constants
(src/run.py:100-201), the dynamic-batch accumulator (:215-244), the occupancy-update cadence
(:248-249), loss + TV regulariser (:252-256) and the un-unscaled GradScaler quirk (:259-260: the loss
is multiplied by 2**10 and never unscaled, so Adam sees gradients x1024).

What is new: rays shard across ranks, every rank accumulates its own ~batch*S packed
samples, the MSE is normalised by the GLOBAL ray count, gradients are summed with NCCL all-reduce, and
the occupancy update is sharded by depth slices and reconciled with an all-gather.  Data parsing, eval
metrics, image writing and the CLI are out of scope (callers own them).
"""
from __future__ import annotations

import ctypes as C
import os
from dataclasses import dataclass
from typing import Dict, List, Tuple

import torch
import torch.distributed as dist

from . import _lib
from .core import (ContractionAABB, ContractionMip360, NerfRenderer, OccupancyGrid, RayMarcherAABB,
                   RayMarcherUnbounded, RayProvider, tag_partition, tag_steps, tagged_steps)
from .optim import FusedAdam
from .models import (CobafaFeatureField, KPlanesFeatureField, VanillaColorDecoder, VanillaFeatureMLP,
                     VanillaOpacityDecoder)


# ---- data-parallel plumbing (device-agnostic; exercised on CPU with gloo in tests/test_dp_gloo.py) ----


def global_ray_count(n_local: int, device, world: int) -> torch.Tensor:
    """Number of rays in the union batch of all ranks (0-dim float tensor on `device`)."""
    n = torch.tensor(float(n_local), device=device)
    if world > 1:
        dist.all_reduce(n)
    return n


def dp_mse(rendered: torch.Tensor, target: torch.Tensor, n_rays_global: torch.Tensor) -> torch.Tensor:
    """This rank's share of MSELoss over the union batch (src/run.py:252): the sum over ranks equals
    mse_loss(cat(rendered), cat(target)) although every rank holds a different number of rays."""
    return ((rendered - target) ** 2).sum() / (n_rays_global * rendered.size(-1))


def _flat_dense(t: torch.Tensor) -> torch.Tensor:
    """1-D view over the memory of a dense tensor of any stride order (the channels-last plane gradients are dense
    but not `contiguous()`, which the collectives require)."""
    if t.is_contiguous():
        return t.view(-1)
    return torch.as_strided(t, (t.numel(),), (1,), t.storage_offset())


def allreduce_gradients(params, world: int) -> None:
    """Sum the per-rank gradients (NCCL over NVLink on the GPU box): ONE collective over a flat copy of all gradients
    instead of one per parameter (35 launches and their latencies for Cobafa), copied back with one foreach op."""
    if world <= 1:
        return
    flats = [_flat_dense(p.grad) for p in params if p.grad is not None]
    if not flats:
        return
    buf = torch.cat(flats)
    dist.all_reduce(buf)
    torch._foreach_copy_(flats, list(buf.split([f.numel() for f in flats])))


def shard_slices(depth: int, rank: int, world: int) -> Tuple[int, int]:
    """Depth slices [z0, z1) of the occupancy grid owned by `rank`."""
    if depth % world != 0:
        raise ValueError("occupancy grid depth must be divisible by the number of ranks")
    per = depth // world
    return rank * per, (rank + 1) * per


class MultiStepLR:
    """torch.optim.lr_scheduler.MultiStepLR as the reference configures it (src/run.py:188-199) without the per-step
    Python bookkeeping of the torch class (~0.2 ms per call, comparable to a whole iteration's device time here):
    lr = base_lr * gamma ** (number of milestones <= epoch), milestones counted with multiplicity."""

    def __init__(self, optimizer, milestones, gamma: float):
        self.optimizer, self.milestones, self.gamma = optimizer, sorted(int(m) for m in milestones), float(gamma)
        self.base_lrs = [g["lr"] for g in optimizer.param_groups]
        self.last_epoch = 0

    def step(self) -> None:
        self.last_epoch += 1
        if self.last_epoch in self.milestones:  # torch multiplies the CURRENT lr by gamma once per occurrence
            k = self.milestones.count(self.last_epoch)
            for g in self.optimizer.param_groups:
                g["lr"] = g["lr"] * self.gamma ** k

    def get_last_lr(self):
        return [g["lr"] for g in self.optimizer.param_groups]


class RayStore:
    """Rays + colours of a scene, shuffled per epoch like DataLoader(shuffle=True) (src/run.py:116-122).
    Lives on `device` (HBM-resident, no per-batch H2D) or in pinned host memory (`host=True`), in which
    case every batch's rows cross the host link inside next(): the reference collates them on the host and copies
    (src/run.py:226-228), here the GPU reads exactly those rows out of the pinned table (tnf_gather_rows, zero-copy).
    With world_size > 1 each rank shuffles its own disjoint 1/world of the rays (ray index % world == rank)."""

    def __init__(self, rays_o: torch.Tensor, rays_d: torch.Tensor, rgbs: torch.Tensor, device, host: bool = False,
                 seed: int = 0, rank: int = 0, world: int = 1):
        self.device, self.host, self.rank, self.world = torch.device(device), host, rank, world
        data = torch.cat([rays_o, rays_d, rgbs], -1).float().contiguous()  # [n, 9]
        self.data = data.pin_memory() if host else data.to(self.device)
        self.n = data.size(0)
        # The order is an incremental Fisher-Yates shuffle on the host (tnf_shuffle_next): each call shuffles just the
        # positions it hands out, so no step ever pays an O(n) randperm at an epoch boundary (tens of milliseconds for
        # millions of rays).  Rank r owns the rays r, r + world, r + 2 world, ... and shuffles that subset on its own (its
        # own seeded generator): the ranks stay disjoint, every ray is handed out once per epoch, and the host work per batch
        # does not grow with the number of ranks.
        self._order = torch.arange(rank, self.n, world, dtype=torch.int64)
        self._m = self._order.numel()
        self._rng = C.c_uint64((0x9E3779B97F4A7C15 * (2 * (int(seed) * 1024 + rank) + 1)) & 0xFFFFFFFFFFFFFFFF)
        self._fresh = C.c_int64(0)     # first global position never drawn so far
        self._pos = 0                  # position of the next batch in this rank's order
        self.h2d_bytes = 0
        self._idx_ring, self._idx_i = [], 0

    def rewind(self, n_rays: int) -> None:
        """Give back the last `n_rays` rays handed out by next() (speculatively marched, not used): the same rays come
        out again, in the same order."""
        self._pos -= n_rays
        assert self._pos >= 0

    def remaining(self) -> int:
        """Rays of this rank left before its order wraps around (the next epoch's swaps start there).  A batch may
        straddle the wrap, but it must not be handed back afterwards: the next epoch's swaps would already have moved
        entries of the positions being replayed."""
        return self._m - (self._pos % self._m)

    def _next_indices(self, batch: int) -> torch.Tensor:
        on_gpu = self.device.type == "cuda"
        if on_gpu:
            # pinned staging for the asynchronous upload of the indices: a ring of four slots, each guarded by the event of
            # its last upload (a slot is rewritten only after that copy has run; in the trainer it always has, because the
            # batch loop reads the sample count back before it draws the next batch)
            slot = self._idx_i
            if len(self._idx_ring) <= slot:   # 1 MB per slot up front: pinning memory is a slow, device-synchronising call
                self._idx_ring.append([torch.empty(max(batch, 1 << 17), dtype=torch.int64).pin_memory(), None])
            buf, ev = self._idx_ring[slot]
            if ev is not None:
                ev.synchronize()
            if buf.numel() < batch:
                buf = self._idx_ring[slot][0] = torch.empty(2 * batch, dtype=torch.int64).pin_memory()
            out = buf[:batch]
            self._idx_i = (slot + 1) % 4
        else:
            out = torch.empty(batch, dtype=torch.int64)
        _lib.check(_lib.load().tnf_shuffle_next(self._order.data_ptr(), self._m, self._pos, batch, 0, 1,
                                                C.byref(self._fresh), C.byref(self._rng), out.data_ptr()), "tnf_shuffle_next")
        self._pos += batch
        self.last_indices = out   # host view of the batch just drawn (valid until the staging slot is reused)
        if not on_gpu:
            return out
        dev_idx = out.to(self.device, non_blocking=True)
        self._idx_ring[slot][1] = torch.cuda.current_stream(self.device).record_event()
        return dev_idx

    def next(self, batch: int) -> Tuple[torch.Tensor, torch.Tensor, torch.Tensor]:
        idx = self._next_indices(batch)
        if self.device.type != "cuda":   # CPU plumbing tests (gloo): plain torch
            rows = torch.index_select(self.data, 0, idx)
            return rows[:, 0:3], rows[:, 3:6], rows[:, 6:9]
        # capacity in 16 Ki-row quanta, like RayProvider.pack: the chunk count of a batch wanders, and a request slightly
        # larger than every cached block makes the caching allocator cudaMalloc in the middle of a step
        cap = (batch + 16383) & ~16383
        rows = torch.empty(cap, 9, device=self.device)[:batch]
        # the GPU picks the rows itself -- from HBM, or (host=True) straight out of the pinned host table over the host link:
        # the H2D copy of exactly this batch's rows, without a host-side gather or a staging buffer (the reference's loader
        # collates on the host and copies, src/run.py:116-122,226-228)
        with torch.cuda.device(self.device):
            _lib.call("tnf_gather_rows", self.data.data_ptr(), self.n, 9, idx.data_ptr(), batch, rows.data_ptr(), _lib.stream_ptr(),
                      nbytes=batch * (36 + 36 + 8))
        if self.host:
            self.h2d_bytes += batch * (36 + 8)   # the rows (zero-copy reads) + their indices
        return rows[:, 0:3], rows[:, 3:6], rows[:, 6:9]


def group_render_chunks(ends: List[int], ray_ends: List[int], max_samples: int) -> List[Tuple[int, int, int, int]]:
    """Chunks of a whole-image render (Trainer.render).  `ends[i]` = packed samples of the image up to and including ray
    block i, `ray_ends[i]` = rays up to and including it.  Consecutive blocks are grouped while the chunk stays within
    `max_samples` packed samples (a block larger than that is a chunk of its own).  -> [(ray0, ray1, sample0, sample1)]."""
    out, r0, s0, i = [], 0, 0, 0
    while i < len(ends):
        j = i
        while j + 1 < len(ends) and ends[j + 1] - s0 <= max_samples:
            j += 1
        out.append((r0, ray_ends[j], s0, ends[j]))
        r0, s0, i = ray_ends[j], ends[j], j + 1
    return out


@dataclass
class TrainConfig:
    """The hot-path subset of the reference's TrainConfig (src/run.py:83-94) + scene description."""
    method: str = "kplanes"             # vanilla | kplanes | cobafa
    scene_type: str = "aabb"            # aabb | unbounded
    batch_size: int = 1024
    n_samples: int = 256
    scene_scale: float = 1.0            # RaysDataset.scene_scale (unbounded marcher range)
    bg_color: Tuple[float, float, float] | None = (1.0, 1.0, 1.0)
    grad_scale: float = 2.0 ** 10       # GradScaler(2**10) that is never unscaled (src/run.py:201,259)
    accumulate: str = "batched"         # "sequential" = the reference's chunk-by-chunk loop (one sync per chunk)
    prefetch: bool = True               # march the coming batches on a side stream while this step trains
    prefetch_depth: int = 2             # batches kept marched ahead (2: the host's per-batch sync never waits for the step
                                        # in flight, so a host hiccup on one rank is absorbed instead of stalling a collective)
    fused_tv_grad: bool = True          # TV gradient written straight into the plane grads (no autograd temporaries)
    fused_step: bool = True             # K-Planes / Cobafa: forward+loss+backward as one C-ABI call sequence (fused.py,
                                        # fused_cobafa.py), no autograd
    max_inflight_steps: int = 3         # the host enqueues at most this many iterations ahead of the GPU (0 = unbounded): a
                                        # free-running host ends up a launch-queue's worth of steps ahead, every batch it has
                                        # marched stays allocated until the GPU gets there, and the allocator's cudaMalloc
                                        # calls then wait for the whole queue
    overlap_plane_adam: bool = False    # single GPU, fused step: Adam of the planes (99.8 % of the parameters, HBM-bound) on a
                                        # second stream beside the heads' weight-gradient kernels.  Measured and OFF: the
                                        # one-CTA-per-SM tensor-core kernels then start late or share HBM badly -- 188.6 M
                                        # samples/s sequential vs 166.9 M (8064 Adam blocks) / 151-164 M (persistent 148/296)
    manual_gc: bool = True              # collect Python garbage at the occupancy-update cadence instead of at random steps
                                        # (a generation-2 pause on ONE rank stalls every rank at the next collective).
                                        # PROCESS-WIDE side effect while the trainer lives: gc.freeze() + gc.disable();
                                        # Trainer.close() / `with Trainer(...)` restores the collector
    occupancy_jitter: str = "device"    # "cpu" = the reference's generator stream
    dp_mode: str = "peer"               # world > 1, fused K-Planes step: "peer" = parameters/gradients in symmetric memory, ONE
                                        # kernel per step reduces the gradients over NVLink, applies Adam to this rank's share
                                        # and writes the new parameters to every rank (dp.PeerMemory, csrc/dp.cu);
                                        # "nccl" = the round-1 baseline: all-reduce of the flat gradient + replicated Adam
    seed: int = 0


class Trainer:
    def __init__(self, cfg: TrainConfig, store: RayStore, device, rank: int = 0, world: int = 1):
        self.cfg, self.store, self.device, self.rank, self.world = cfg, store, torch.device(device), rank, world
        bs_ratio = 4096 / cfg.batch_size
        self.steps = int(2048 * bs_ratio)
        self.occupancy_grid_updates = int(16 * bs_ratio)
        thr, res = 0.01, 128
        decay = thr ** (1 / 16)
        dev = self.device
        if cfg.method == "vanilla":
            feature_module = VanillaFeatureMLP(10, 256, 8)
        elif cfg.method == "kplanes":
            feature_module = KPlanesFeatureField(32)
        elif cfg.method == "cobafa":
            feature_module = CobafaFeatureField(basis_res=torch.linspace(32.0, 128, 6).int().tolist(), coef_res=64,
                                                freqs=torch.linspace(2.0, 8.0, 6).tolist(), channels=[8, 8, 8, 4, 4, 4],
                                                mlp_hidden_dim=128)
        else:
            raise NotImplementedError(f"Unknown method {cfg.method}.")
        dim = feature_module.feature_dim
        sigma_decoder = VanillaOpacityDecoder(dim)
        rgb_decoder = VanillaColorDecoder(8, dim, 64, 3)
        if cfg.scene_type == "unbounded":
            self.ray_marcher = RayMarcherUnbounded(cfg.n_samples, 0.1, 1e5, uniform_range=cfg.scene_scale)
            contraction = ContractionMip360(order=float("inf"))
        elif cfg.scene_type == "aabb":
            aabb = torch.tensor([[-1.5, -1.5, -1.5], [1.5, 1.5, 1.5]]).to(dev)
            self.ray_marcher = RayMarcherAABB(aabb, cfg.n_samples, 0.1)
            contraction = ContractionAABB(aabb)
        else:
            raise NotImplementedError(f"Unknown scene type {cfg.scene_type}.")
        self.occupancy_grid = OccupancyGrid(size=res, step_size=self.ray_marcher.step_size, threshold=thr,
                                            decay=decay).to(dev)
        self.occupancy_grid.jitter_source = cfg.occupancy_jitter
        # 16 depth slices (2^18 points) per sigma_fn call: the update's temporaries stay the size of a training batch's
        # (no GB-sized allocations that the caching allocator would have to cudaMalloc around live per-step buffers)
        self.occupancy_grid.slices_per_call = 16
        self.ray_provider = RayProvider(self.occupancy_grid, contraction, self.ray_marcher)
        bg = None if cfg.bg_color is None else torch.tensor(cfg.bg_color)
        self.renderer = NerfRenderer(feature_module, sigma_decoder, rgb_decoder, bg_color=bg).to(dev)
        if world > 1:  # identical replicas: broadcast rank 0's initial parameters
            for p in self.renderer.parameters():
                dist.broadcast(_flat_dense(p.data), 0)
        self.optimizer = FusedAdam(self.renderer.parameters(), lr=1e-2, eps=1e-15, weight_decay=1e-5)
        s = self.steps
        self.scheduler = MultiStepLR(self.optimizer, milestones=[s // 2, s * 3 // 4, s * 5 // 6, s * 9 // 10], gamma=0.33)
        self.tv_reg_alpha, self.l1_reg_alpha = 0.0001, 0.0
        self.train_step = 0
        self._fused = None
        if cfg.fused_step and cfg.method == "kplanes" and self.device.type == "cuda" and self.l1_reg_alpha == 0.0:
            from .fused import FusedKPlanesStep
            if FusedKPlanesStep.supported(self.renderer):
                if cfg.dp_mode not in ("peer", "nccl"):
                    raise ValueError(f"unknown dp_mode {cfg.dp_mode!r}")
                self._fused = FusedKPlanesStep(self.renderer, tv_alpha=self.tv_reg_alpha, grad_scale=cfg.grad_scale,
                                               world=world, rank=rank, peer_update=(world > 1 and cfg.dp_mode == "peer"))
                if self._fused.peer is not None:
                    # optimizer.state keeps torch.optim.Adam's keys (checkpoints): exp_avg / exp_avg_sq are views of the flat
                    # state the peer kernel updates; only this rank's share is live until PeerMemory.gather_optimizer_state()
                    peer, off = self._fused.peer, 0
                    for p in self._fused.params:
                        view = lambda flat: torch.as_strided(flat, p.shape, p.stride(), off)
                        self.optimizer.state[p] = {"step": 0, "exp_avg": view(peer.exp_avg), "exp_avg_sq": view(peer.exp_avg_sq)}
                        off += (p.numel() + 3) // 4 * 4
        # Cobafa: the same idea (fused_cobafa.py) -- the module path's host work exceeds the iteration's kernel time
        self._fused_cobafa = None
        if cfg.fused_step and cfg.method == "cobafa" and self.device.type == "cuda":
            from .fused_cobafa import FusedCobafaStep
            if FusedCobafaStep.supported(self.renderer):
                self._fused_cobafa = FusedCobafaStep(self.renderer, grad_scale=cfg.grad_scale, world=world)
        self._chunks_guess = 0.0
        if world > 1 and self.device.type == "cuda":
            self._warm_collectives()
        # The coming batches are marched on a HIGH-PRIORITY stream: their short kernels are then not queued behind the main
        # stream's long ones while the host waits for the sample count (measured: 186.7 -> 192.6 M samples/s, end to end
        # 188.6 -> 192.0 M, and a much tighter per-step distribution; TNF_SIDE_PRIORITY=0 restores equal priorities)
        prio = -1 if os.environ.get("TNF_SIDE_PRIORITY", "1") == "1" else 0
        self._side = torch.cuda.Stream(device=self.device, priority=prio) if (cfg.prefetch and self.device.type == "cuda") else None
        if self.device.type == "cuda":
            self._prewarm_allocator()
        self.post_update = None         # optional callable(trainer) run right after every occupancy update
        self._gc_frozen = False
        self._queue: List = []          # prefetched (batch, done event), oldest first
        self._inflight: List = []       # end-of-iteration events of the iterations enqueued and not yet waited for
        self._adam_blocks = int(os.environ.get("TNF_ADAM_BLOCKS", "0")) or (
            torch.cuda.get_device_properties(self.device).multi_processor_count if self.device.type == "cuda" else 0)
        self._opt_stream = None         # second stream for the planes' optimiser update (overlap_plane_adam)
        self._loss_ring = None          # pinned host ring the loss of every iteration is copied into (read_loss)
        self._grid_event = None         # recorded after the latest occupancy update: batches marched later must see it
        self.last: Dict[str, float] = {}

    def _prewarm_allocator(self, blocks: int = 24, mb: int = 64) -> None:
        """Leave ~1.5 GB of free blocks in PyTorch's caching allocator.  The per-batch buffers (rays, jitter, bitfield,
        packed rows) change size with the dynamic batch; a size the cache has not seen makes the allocator call cudaMalloc
        in the middle of a step, which waits for all queued GPU work and -- in a process with peer access enabled (NCCL,
        symmetric memory) -- maps the new memory into every peer: stalls of 5-60 ms were measured inside 64-step windows.
        Free 64 MB blocks are split on demand instead."""
        if os.environ.get("TNF_PREWARM_ALLOCATOR", "1") == "0":
            return
        # The allocator keeps separate free lists PER STREAM: the batches are allocated on the prefetch stream, the
        # iteration's temporaries on the main stream -- both get their own reserve.  Large blocks (split on demand) and the
        # SMALL pool (requests < 1 MB are carved out of 2 MB segments: ray indices, packing info, bitfields, counters): a new
        # 2 MB segment in the middle of a run was measured to stall one step for 82 ms.
        side = getattr(self, "_side", None)
        for stream in ([torch.cuda.current_stream(self.device)] + ([side] if side is not None else [])):
            with torch.cuda.stream(stream):
                hold = [torch.empty(mb << 20, dtype=torch.uint8, device=self.device) for _ in range(blocks if stream is not side else blocks // 2)]
                hold += [torch.empty(512 << 10, dtype=torch.uint8, device=self.device) for _ in range(128)]
                del hold

    def _warm_collectives(self) -> None:
        """Run every collective of an iteration once at its real size (gradient all-reduce, ray-count all-reduce,
        occupancy all-gather) so NCCL's lazy channel/buffer set-up happens at construction, not inside a step."""
        peer = self._fused is not None and self._fused.peer is not None
        grads = self._fused.flat_grad if self._fused is not None else torch.zeros(
            sum(p.numel() for p in self.renderer.parameters()), device=self.device)
        for _ in range(2):
            if not peer:
                dist.all_reduce(grads)
            global_ray_count(1, self.device, self.world)
            g = self.occupancy_grid.grid
            z0, z1 = shard_slices(g.size(0), self.rank, self.world)
            dist.all_gather_into_tensor(torch.empty_like(g).view(-1), g[z0:z1].reshape(-1).clone())
        if self._fused is not None and not peer:
            grads.zero_()
        torch.cuda.synchronize(self.device)

    # ---- a11: dynamic batch accumulator (src/run.py:215-244) -----------------------------------
    @torch.no_grad()
    def next_batch(self):
        if self.cfg.accumulate == "sequential":
            return self._next_batch_sequential()
        return self._next_batch_batched()

    @torch.no_grad()
    def _next_batch_sequential(self):
        """The reference's loop verbatim: one provider call (and host sync) per chunk of batch_size rays."""
        target = self.cfg.batch_size * self.cfg.n_samples
        current, projected, k = 0, 0, 0
        acc_s, acc_i, acc_rgb, acc_steps = [], [], [], []
        while projected < target:
            rays_o, rays_d, rgbs = self.store.next(self.cfg.batch_size)
            samples, info = self.ray_provider(rays_o, rays_d, training=True, info_offset=current)
            acc_s.append(samples)
            acc_steps.append(tagged_steps(samples))
            acc_i.append(info)
            acc_rgb.append(rgbs)
            current += samples.size(0)
            k += 1
            projected = int(current * (1 + 1 / k))
            if k > 4096:
                raise RuntimeError("occupancy grid rejects every sample: cannot fill a batch")
        packed = torch.cat(acc_s, 0)
        tag_steps(packed, torch.cat(acc_steps, 0))
        info = torch.cat(acc_i, 0)
        tag_partition(info)  # consecutive batches were packed at consecutive offsets
        return packed, torch.cat(acc_rgb, 0), info

    @torch.no_grad()
    def _next_batch_batched(self):
        """Same rays, same jitter stream, same packed output as the sequential loop, but the lattice of several
        chunks is marched in ONE launch and the chunk count k is then chosen on the host by the reference's rule
        (`int(cur*(1+1/k)) >= target`) from the per-chunk totals: one host sync per step instead of one per chunk.
        Chunks marched speculatively beyond k are handed back (rays rewound, generator offset restored)."""
        B, S = self.cfg.batch_size, self.cfg.n_samples
        target = B * S
        dev = self.device
        gen = torch.cuda.default_generators[dev.index if dev.index is not None else torch.cuda.current_device()]
        current, k_done = 0, 0
        parts = []
        while True:
            K = max(1, int(self._chunks_guess) + 1)
            rem = self.store.remaining()
            # never speculate across the epoch boundary (RayStore.rewind must not cross it): cap the chunks marched at
            # once at the rays left in the epoch; a chunk that straddles the wrap is marched alone (K = 1 is never rewound)
            K = min(K, rem // B) if rem >= B else 1
            rays_o, rays_d, rgbs = self.store.next(K * B)
            noise = torch.empty(((K + 3) & ~3) * B, S, device=dev)[:K * B]   # allocation quantised to 4 chunks (see RayStore.next)
            offsets = []
            for i in range(K):  # chunk-wise draws: the same generator consumption as the reference's rand_like per chunk
                torch.rand(B, S, out=noise[i * B:(i + 1) * B])
                offsets.append(gen.get_offset())
            h = self.ray_provider.count(rays_o, rays_d, True, noise, info_offset=current)
            ends = (h["info"][B - 1::B, 0].long() + h["info"][B - 1::B, 1].long()).tolist()  # the one host sync
            k_use = None
            for i, e in enumerate(ends):
                k_tot = k_done + i + 1
                if int(e * (1 + 1 / k_tot)) >= target:
                    k_use = i + 1
                    break
            used = K if k_use is None else k_use
            n_here = ends[used - 1] - current
            packed, info = self.ray_provider.pack(h, used * B, n_here)
            parts.append((packed, info, rgbs[:used * B]))
            current = ends[used - 1]
            k_done += used
            if k_use is not None:
                if used < K:
                    self.store.rewind((K - used) * B)
                    gen.set_offset(offsets[used - 1])
                break
            self._chunks_guess = max(self._chunks_guess, k_done) * 1.5
            if k_done > 4096:
                raise RuntimeError("occupancy grid rejects every sample: cannot fill a batch")
        self._chunks_guess = 0.5 * self._chunks_guess + 0.5 * k_done if self._chunks_guess else float(k_done)
        if len(parts) == 1:
            packed, info, rgb = parts[0]
        else:
            packed = torch.cat([p[0] for p in parts], 0)
            tag_steps(packed, torch.cat([tagged_steps(p[0]) for p in parts], 0))
            info = torch.cat([p[1] for p in parts], 0)
            rgb = torch.cat([p[2] for p in parts], 0)
        tag_partition(info)
        if self._fused is not None:
            self._fused.sort_batch(packed)   # per-orientation counting sort for the sorted plane scatter (same stream as the pack)
        return packed, rgb, info

    # ---- occupancy update, sharded by depth slice across ranks ----------------------------------
    @torch.no_grad()
    def update_occupancy(self):
        og = self.occupancy_grid
        if self._fused is not None:
            sigma_fn = self._fused.density   # same kernels on the iteration's workspaces: no allocations in the update
        elif getattr(self, "_fused_cobafa", None) is not None:
            sigma_fn = self._fused_cobafa.density
        else:
            sigma_fn = lambda t: self.renderer.sigma_decoder(self.renderer.feature_module(t))
        if self.world == 1:
            og.update(sigma_fn)
            return
        z0, z1 = shard_slices(og.grid.size(0), self.rank, self.world)
        og.update_slices(sigma_fn, z0, z1)
        dist.all_gather_into_tensor(og.grid.view(-1), og.grid[z0:z1].reshape(-1).clone())
        og._set_mean_from_grid()

    # ---- one training iteration (src/run.py:246-261) -------------------------------------------
    def _prefetch(self) -> None:
        """Keep `prefetch_depth` batches marched ahead on the side stream.  A batch depends only on the occupancy grid:
        the batch of iteration j is generated before j's own update (src/run.py:215-249), so it must see the updates of
        iterations < j.  After enqueueing iteration t, batch t+1 is always eligible; batch t+2 only if iteration t+1 does
        not update the grid.  Batches are marched in order, so rays and jitter are consumed in the reference's order."""
        side = self._side
        t = self.train_step  # index of the next iteration to run
        while len(self._queue) < max(1, self.cfg.prefetch_depth):
            j = t + len(self._queue)          # iteration this batch is for
            if any(i % self.occupancy_grid_updates == 0 for i in range(t, j)):
                break                          # an update between now and j has not been enqueued yet
            if self._grid_event is not None:
                side.wait_event(self._grid_event)
            with torch.cuda.stream(side):
                batch = self.next_batch()
                done = side.record_event()
            self._queue.append((batch, done))

    def _take_batch(self):
        if not self._queue:
            return self.next_batch()
        (packed, rgbs, info), done = self._queue.pop(0)
        main = torch.cuda.current_stream(self.device)
        main.wait_event(done)
        extra = getattr(packed, "_tnf_ksort", None)
        for t in (packed, tagged_steps(packed), rgbs, info) + (tuple(extra[:2]) if extra else ()):
            t.record_stream(main)
        return packed, rgbs, info

    def step(self) -> Dict[str, float]:
        if self.cfg.manual_gc and self.device.type == "cuda":
            import gc
            if not self._gc_frozen:
                self._gc_was_enabled = gc.isenabled()
                gc.collect()
                gc.freeze()    # everything allocated so far (torch's ~10^6 module objects) leaves the collector's reach:
                gc.disable()   # later collections only scan what the loop itself created (undone by close())
                self._gc_frozen = True
            elif self.train_step % self.occupancy_grid_updates == 0:
                gc.collect()
        if self.cfg.max_inflight_steps > 0 and self.device.type == "cuda":
            while len(self._inflight) >= self.cfg.max_inflight_steps:
                self._inflight.pop(0).synchronize()
        packed, rgbs, info = self._take_batch()
        if not self.renderer.training:
            self.renderer.train()
        if self.train_step % self.occupancy_grid_updates == 0:
            self.update_occupancy()
            if self.post_update is not None:
                self.post_update(self)   # e.g. the benchmark pins the occupancy state here, inside the stream order
            if self._side is not None:
                self._grid_event = torch.cuda.current_stream(self.device).record_event()
        if self._fused is not None:
            out = self._step_fused(packed, rgbs, info)
            self._after_first_steps()
            return out
        if self._fused_cobafa is not None:
            return self._step_fused_cobafa(packed, rgbs, info)
        rendered = self.renderer(packed, info)
        loss = dp_mse(rendered, rgbs, global_ray_count(info.size(0), self.device, self.world))
        tv_direct = 0.0
        if self.cfg.method == "kplanes":
            fm = self.renderer.feature_module
            if self.cfg.fused_tv_grad:
                # loss += tv_alpha * loss_tv (src/run.py:254-255): the value joins the reported loss, its gradient
                # (a constant-coefficient stencil) is added straight into the plane gradients after backward()
                with torch.no_grad():
                    loss_report = loss.detach() + fm.loss_tv() * (self.tv_reg_alpha / self.world)  # type: ignore
                tv_direct = self.tv_reg_alpha / self.world * self.cfg.grad_scale
            else:
                reg = fm.loss_tv() * self.tv_reg_alpha  # type: ignore
                loss = loss + reg / self.world  # summed over ranks it counts once
            if self.l1_reg_alpha != 0.0:  # the reference multiplies by 0. (src/run.py:114,256): no contribution
                loss = loss + fm.loss_l1() * (self.l1_reg_alpha / self.world)  # type: ignore
        self.optimizer.zero_grad()
        (loss * self.cfg.grad_scale).backward()
        if tv_direct:
            self.renderer.feature_module.add_tv_grad_(tv_direct)  # type: ignore
            loss = loss_report
        allreduce_gradients(self.renderer.parameters(), self.world)
        self.optimizer.step()
        self.scheduler.step()
        self.train_step += 1
        self._publish_loss(loss)
        self._mark_enqueued()
        if self._side is not None:
            # everything above is enqueued; the coming batches are marched on the side stream while the GPU is still
            # busy with this step's backward + optimiser, so their host sync no longer stalls the step
            self._prefetch()
        self.last = {"loss": loss.detach(), "n_samples": packed.size(0), "n_rays": info.size(0)}
        return self.last

    # ---- process-wide side effects are undone here -------------------------------------------------
    def close(self) -> None:
        """Restore what step() changed process-wide: with `manual_gc` the cyclic collector is frozen and disabled while
        the trainer runs (collections happen at the occupancy-update cadence); close() -- also called by `with Trainer(...)`
        and on garbage collection of the trainer -- unfreezes it and re-enables it if it was enabled before."""
        if getattr(self, "_fused", None) is not None and self.device.type == "cuda":
            self._fused.wait_updates()   # peer-update mode: the side stream's last parameter updates join the main stream
        if self._gc_frozen:
            import gc
            gc.unfreeze()
            if getattr(self, "_gc_was_enabled", True):
                gc.enable()
            self._gc_frozen = False

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()
        return False

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _mark_enqueued(self) -> None:
        if self.cfg.max_inflight_steps > 0 and self.device.type == "cuda":
            self._inflight.append(torch.cuda.current_stream(self.device).record_event())

    # ---- loss read-back without stalling the pipeline (the reference reads loss.item() every step, src/run.py:263) ----
    _LOSS_RING = 8

    def _publish_loss(self, loss: torch.Tensor) -> None:
        """Device -> pinned-host copy of this iteration's loss, asynchronous on the main stream."""
        if self.device.type != "cuda":
            return
        if self._loss_ring is None:
            self._loss_ring = (torch.zeros(self._LOSS_RING, dtype=torch.float32).pin_memory(),
                               [torch.cuda.Event() for _ in range(self._LOSS_RING)])
        buf, evs = self._loss_ring
        slot = (self.train_step - 1) % self._LOSS_RING
        buf[slot:slot + 1].copy_(loss.detach().reshape(1).float(), non_blocking=True)
        evs[slot].record(torch.cuda.current_stream(self.device))

    def read_loss(self, step: int | None = None) -> float:
        """Host value of the loss of iteration `step` (0-based; default: the latest).  Waits only for that iteration's
        copy, so reading iteration t-1 while t is in flight costs nothing; at most _LOSS_RING iterations are kept."""
        last = self.train_step - 1
        step = last if step is None else step
        if self._loss_ring is None or not (max(0, last - self._LOSS_RING + 1) <= step <= last):
            raise ValueError(f"loss of iteration {step} is not available (latest: {last})")
        buf, evs = self._loss_ring
        evs[step % self._LOSS_RING].synchronize()
        return float(buf[step % self._LOSS_RING])

    def _step_fused(self, packed, rgbs, info) -> Dict[str, float]:
        """Same iteration through fused.FusedKPlanesStep: identical kernels and order, no autograd / glue ops."""
        if self._fused.peer is not None:
            return self._step_fused_peer(packed, rgbs, info)
        n_glob = work = None
        if self.world > 1:  # ray count of the union batch: reduced while the forward runs
            n_glob = torch.tensor(float(info.size(0)), device=self.device)
            work = dist.all_reduce(n_glob, async_op=True)
        hook = None
        if self.cfg.overlap_plane_adam and self.world == 1:
            main = torch.cuda.current_stream(self.device)
            if self._opt_stream is None:
                self._opt_stream = torch.cuda.Stream(device=self.device)
            planes = self._fused.planes

            def hook():
                self._opt_stream.wait_event(main.record_event())
                with torch.cuda.stream(self._opt_stream):
                    self.optimizer.step(subset=planes, max_blocks=self._adam_blocks)
                    self._planes_done = self._opt_stream.record_event()
        out = self._fused.forward_backward(packed, info, rgbs, n_glob, reduce=self.world > 1, n_rays_work=work,
                                           after_plane_grads=hook)  # incl. the gradient all-reduce
        if hook is not None:
            plane_ids = {id(p) for p in self._fused.planes}
            self.optimizer.step(subset=[p for p in self._fused.params if id(p) not in plane_ids])
            torch.cuda.current_stream(self.device).wait_event(self._planes_done)
        else:
            self.optimizer.step()
        self.scheduler.step()
        self.train_step += 1
        self._publish_loss(out["loss"])
        self._mark_enqueued()
        if self._side is not None:
            self._prefetch()
        self.last = {"loss": out["loss"], "n_samples": packed.size(0), "n_rays": info.size(0)}
        return self.last

    def _step_fused_cobafa(self, packed, rgbs, info) -> Dict[str, float]:
        """The Cobafa iteration through fused_cobafa.FusedCobafaStep: same kernels and order as the module path, no autograd /
        glue ops; data-parallel: the union batch's ray count is reduced while the forward runs, the flat gradient in one
        collective, then the replicated Adam (src/run.py:251-261)."""
        n_glob = work = None
        if self.world > 1:
            n_glob = torch.tensor(float(info.size(0)), device=self.device)
            work = dist.all_reduce(n_glob, async_op=True)
        out = self._fused_cobafa.forward_backward(packed, info, rgbs, n_glob, reduce=self.world > 1, n_rays_work=work)
        self.optimizer.step()
        self.scheduler.step()
        self.train_step += 1
        self._publish_loss(out["loss"])
        self._mark_enqueued()
        if self._side is not None:
            self._prefetch()
        self.last = {"loss": out["loss"], "n_samples": packed.size(0), "n_rays": info.size(0)}
        return self.last

    def _step_fused_peer(self, packed, rgbs, info) -> Dict[str, float]:
        """Data-parallel iteration with the update over NVLink peer memory: no NCCL call in the step.  The ray count goes to
        the peers' slot tables before the forward; forward_backward consumes the union count, and issues the planes' and the
        heads' reduce + Adam + broadcast kernels (src/run.py:258-261's optimizer.step(), distributed)."""
        peer = self._fused.peer
        t = self._peer_adam_step = getattr(self, "_peer_adam_step", 0) + 1   # Adam's own 1-based step count
        with torch.cuda.device(self.device):
            if self._fused.peer_overlap:   # nothing on the main stream depends on it: beside the TV pass
                with torch.cuda.stream(self._fused._peer_stream):
                    peer.publish_count(t, info.size(0))
            else:
                peer.publish_count(t, info.size(0))
        g = self.optimizer.param_groups[0]
        out = self._fused.forward_backward(packed, info, rgbs, peer_step={
            "step": t, "lr": g["lr"], "betas": g["betas"], "eps": g["eps"], "weight_decay": g["weight_decay"]})
        for st in self.optimizer.state.values():
            st["step"] = t
        if t % self.occupancy_grid_updates == 0:
            peer.check()
        self.scheduler.step()
        self.train_step += 1
        self._publish_loss(out["loss"])
        self._mark_enqueued()
        if self._side is not None:
            self._prefetch()
        self.last = {"loss": out["loss"], "n_samples": packed.size(0), "n_rays": info.size(0)}
        return self.last

    def _after_first_steps(self) -> None:
        """The iteration's workspaces are allocated lazily by the first steps and split the pre-warmed blocks; top the
        allocator's free lists up again once they exist (steps 1 and 3, and after the first occupancy update)."""
        if self.device.type == "cuda" and self.train_step in (1, 3, self.occupancy_grid_updates + 1):
            self._prewarm_allocator()

    # ---- render half of the path (src/run.py:15-50, without image IO) ---------------------------
    @torch.no_grad()
    def render(self, rays_o: torch.Tensor, rays_d: torch.Tensor, batch_size: int = 2048,
               max_samples: int = 1 << 20, out: torch.Tensor | None = None) -> torch.Tensor:
        """Rendered colours [R,3] (on the device) of the rays of an image, row for row what the reference's loop returns
        (src/run.py:34-44: chunks of `batch_size` rays through ray_provider(training=False) + renderer, concatenated).
        A ray's colour does not depend on which rays share its chunk, so the chunking is free to change:
          * the sample lattice of ALL rays is marched in one launch (keep-mask bitfield + packing info on the device);
          * ONE host sync per image reads the packed-sample total at every `batch_size`-th ray; the host then groups those
            blocks into chunks of at most `max_samples` packed samples (a fixed-capacity buffer pair, reused by every chunk);
          * each chunk is packed straight from the bitfield (no second march, no sync) and rendered forward-only:
            K-Planes through FusedKPlanesStep.render (5 launches, no activations saved), other fields through the modules.
        The reference syncs twice per chunk and copies every chunk to the host; here the image stays on the device until
        the caller reads it."""
        self.renderer.eval()
        dev = self.device
        o = rays_o.reshape(-1, 3).to(dev, torch.float32, non_blocking=True).contiguous()
        d = rays_d.reshape(-1, 3).to(dev, torch.float32, non_blocking=True).contiguous()
        R = o.size(0)
        if out is None:
            out = torch.empty(R, 3, device=dev)
        if R == 0:
            return out
        h = self.ray_provider.count(o, d, training=False)
        info = h["info"]
        idx = list(range(batch_size - 1, R, batch_size))
        if not idx or idx[-1] != R - 1:
            idx.append(R - 1)
        last = torch.tensor(idx, device=dev)
        ends = (info[last, 0].long() + info[last, 1].long()).tolist()      # the one host sync of the image
        ray_ends = [min(R, (i + 1) * batch_size) for i in range(len(ends))]
        cap = max(max_samples, max(e - s for s, e in zip([0] + ends[:-1], ends)))
        cap = (cap + 16383) & ~16383
        if getattr(self, "_render_cap", 0) < cap:
            self._render_buf = (torch.empty(cap, 7, device=dev), torch.empty(cap, device=dev))
            self._render_cap = cap
        pbuf, sbuf = self._render_buf
        bg = self.renderer.bg_color
        for r0, r1, s0, s1 in group_render_chunks(ends, ray_ends, max_samples):
            n = s1 - s0
            if n == 0:   # every sample masked: the reference composites nothing, the rays show the background (src/core.py:251-265)
                out[r0:r1] = 0.0 if bg is None else bg.to(dev).reshape(1, 3)
            else:
                packed, ic = self.ray_provider.pack_range(h, r0, r1, s0, n, pbuf, sbuf)
                if self._fused is not None and self._fused.fused_heads and self._fused.split_xc:
                    self._fused.render(packed, ic, out[r0:r1])
                else:
                    out[r0:r1] = self.renderer(packed, ic)
        return out


@torch.no_grad()
def infer(trainer: "Trainer", poses, batch_size: int = 2048):
    """The reference's `infer` (src/run.py:15-50) without the PNG writing: `poses` yields (rays_o [H,W,3], rays_d [H,W,3])
    per image; returns the list of rendered [H,W,3] images on the host."""
    rendered = []
    for rays_o, rays_d in poses:
        img = trainer.render(rays_o, rays_d, batch_size=batch_size)
        rendered.append(img.view(*rays_o.shape[:-1], 3).cpu())
    return rendered
