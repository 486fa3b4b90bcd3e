"""Data-parallel parameter path over NVLink peer memory (SURVEY section 8e + 8f rank 1).

`PeerMemory` is the plumbing: ONE symmetric-memory allocation per rank (torch.distributed._symmetric_memory: the same
layout on every GPU, mapped into every peer, bound to an NVSwitch multicast address when the fabric has one) holding

    [ flat parameters | flat gradients | barrier flag pads | ray-count slot table ]

and the per-step calls of csrc/dp.cu on it:

    publish_count / sum_counts   the union batch's ray count (MSE normaliser, src/run.py:252) without a collective kernel
    reduce_adam_bcast            gradient sum over the ranks + Adam on this rank's slice + new parameters to every rank,
                                 one kernel (reduce-scatter + sharded optimiser + all-gather of src/run.py:258-261's update)

There is no fallback to NCCL in here: if symmetric memory cannot be set up the constructor raises; the caller chooses the
collective strategy explicitly (run.TrainConfig.dp_mode).
"""
from __future__ import annotations

import ctypes as C
import os
from typing import Tuple

import torch
import torch.distributed as dist

from . import _lib

MAX_RANKS = 16      # TNF_DP_MAX_RANKS
COUNT_SLOTS = 4     # TNF_DP_COUNT_SLOTS
FLAG_PADS = 2       # one per concurrent user of the barrier (planes launch, heads launch)


def slice_of(lo: int, hi: int, rank: int, world: int) -> Tuple[int, int]:
    """Rank `rank`'s share of the element range [lo, hi): contiguous, 4-element aligned, disjoint, covering."""
    n4 = (hi - lo + 3) // 4
    per = (n4 + world - 1) // world
    a = min(hi, lo + 4 * per * rank)
    b = min(hi, lo + 4 * per * (rank + 1))
    return a, b


class PeerMemory:
    def __init__(self, n_params: int, device, rank: int, world: int, n_ctas: int | None = None, use_multicast: bool | None = None):
        import torch.distributed._symmetric_memory as symm_mem
        _lib.load()
        if world > MAX_RANKS:
            raise RuntimeError(f"at most {MAX_RANKS} ranks")
        self.device, self.rank, self.world = torch.device(device), rank, world
        sms = torch.cuda.get_device_properties(self.device).multi_processor_count
        self._n_ctas_arg = int(n_ctas or os.environ.get("TNF_DP_CTAS", 0) or 0)
        self.n = (n_params + 4 * world - 1) // (4 * world) * (4 * world)
        flag_words = FLAG_PADS * 4 * sms * MAX_RANKS             # uint32 [pad][cta][source rank], up to 4 CTAs per SM
        slot_words = 2 * COUNT_SLOTS * MAX_RANKS                  # uint64 [slot][source rank] as pairs of 32-bit words
        self._flag_off = 2 * self.n
        self._slot_off = self._flag_off + flag_words
        total = self._slot_off + slot_words
        group = dist.group.WORLD
        try:
            symm_mem.enable_symm_mem_for_group(group.group_name)
        except Exception:
            pass   # newer releases enable every group implicitly
        self.buf = symm_mem.empty(total, dtype=torch.float32, device=self.device)
        self.buf.zero_()
        self.handle = symm_mem.rendezvous(self.buf, group)
        torch.cuda.synchronize(self.device)
        dist.barrier()
        ptrs = [int(p) for p in self.handle.buffer_ptrs]
        if len(ptrs) != world or ptrs[rank] != self.buf.data_ptr():
            raise RuntimeError("symmetric memory rendezvous returned an unexpected peer table")
        mc = int(getattr(self.handle, "multicast_ptr", 0) or 0)
        if use_multicast is None:
            use_multicast = os.environ.get("TNF_DP_MULTICAST", "1") != "0"
        self.multicast = bool(mc) and use_multicast
        # Grid (measured at 2 ranks, scripts/dp_kernel_bench.py + bench.py): the multicast path is switch-bound -- 372 us for
        # the 132 MB of planes with half a CTA per SM, no faster with more -- and a small grid leaves the SMs to the kernels
        # it runs beside (step 1.43 ms against 1.56 ms for the P2P form, which needs 2 CTAs per SM in flight to cover the
        # NVLink latency: 277 us alone, but it then starves the weight-gradient kernels).
        self.n_ctas = self._n_ctas_arg or (sms // 2 if self.multicast else 2 * sms)
        self.param = self.buf[:self.n]
        self.grad = self.buf[self.n:2 * self.n]
        tab = lambda off: (C.c_void_p * world)(*[p + 4 * off for p in ptrs])
        self._peer_param, self._peer_grad = tab(0), tab(self.n)
        self._peer_flags = [tab(self._flag_off + i * 4 * sms * MAX_RANKS) for i in range(FLAG_PADS)]
        self._peer_slots = tab(self._slot_off)
        self._mc_param = mc if self.multicast else None
        self._mc_grad = mc + 4 * self.n if self.multicast else None
        self._slots_local = ptrs[rank] + 4 * self._slot_off
        self.exp_avg = torch.zeros(self.n, device=self.device)
        self.exp_avg_sq = torch.zeros(self.n, device=self.device)
        self.error = torch.zeros(1, dtype=torch.int32, device=self.device)
        self.n_rays_global = torch.zeros(1, device=self.device)
        self._epoch = [1] * FLAG_PADS
        self._regions = set()   # (lo, hi) ranges updated so far: this rank owns slice_of(lo, hi) of each
        self._err_host = torch.zeros(1, dtype=torch.int32).pin_memory()
        self._err_event = None

    # ---- ray count of the union batch --------------------------------------------------------------
    def publish_count(self, step: int, n_rays: int) -> None:
        """This rank's ray count of iteration `step` (1-based) into every rank's slot table (current stream)."""
        _lib.call("tnf_dp_publish_count", self._peer_slots, self.world, self.rank, step % COUNT_SLOTS, step & 0xFFFFFFFF,
                  float(n_rays), _lib.stream_ptr())

    def sum_counts(self, step: int) -> torch.Tensor:
        """Union-batch ray count of iteration `step` -> the device float the loss kernel reads (current stream)."""
        _lib.call("tnf_dp_sum_counts", self._slots_local, self.world, step % COUNT_SLOTS, step & 0xFFFFFFFF,
                  self.n_rays_global.data_ptr(), self.error.data_ptr(), _lib.stream_ptr())
        return self.n_rays_global

    # ---- the update -------------------------------------------------------------------------------
    def reduce_adam_bcast(self, lo: int, hi: int, pad: int, step: int, lr: float, betas, eps: float, weight_decay: float,
                          stream: int | None = None, n_ctas: int | None = None) -> None:
        """Sum the gradients of [lo, hi) over the ranks, apply Adam to this rank's share of it and write the new parameters
        of that share to every rank.  Collective: every rank calls it with the same [lo, hi), pad and step.
        pad: which barrier flag pad this launch uses (launches that may be in flight together need different pads)."""
        a, b = slice_of(lo, hi, self.rank, self.world)
        self._regions.add((lo, hi))
        epoch = self._epoch[pad]
        self._epoch[pad] = (epoch + 2) & 0xFFFFFFFF
        _lib.call("tnf_dp_reduce_adam_bcast", self._peer_grad, self._peer_param, self._mc_grad, self._mc_param,
                  self.exp_avg.data_ptr(), self.exp_avg_sq.data_ptr(), a, b, self._peer_flags[pad], int(n_ctas or self.n_ctas),
                  self.rank, self.world, epoch, self.error.data_ptr(), float(lr), float(betas[0]), float(betas[1]), float(eps),
                  float(weight_decay), int(step), _lib.stream_ptr() if stream is None else stream,
                  nbytes=(b - a) * (24 + 4 * self.world) + (hi - lo) * 4,
                  label="tnf_dp_reduce_adam_bcast" + ("" if hi - lo > (1 << 20) else "(heads)"))

    def check(self) -> None:
        """Raise if a rank barrier timed out.  Reads the error word of an EARLIER call asynchronously (no pipeline stall):
        the copy queued by the previous check() is inspected, then a new one is queued."""
        if self._err_event is not None:
            self._err_event.synchronize()
            if int(self._err_host[0]) != 0:
                raise RuntimeError(f"data-parallel barrier timed out (code {int(self._err_host[0])}): a peer rank stopped")
        self._err_host.copy_(self.error, non_blocking=True)
        self._err_event = torch.cuda.current_stream(self.device).record_event()

    def gather_optimizer_state(self) -> None:
        """exp_avg / exp_avg_sq exist only on the rank that owns a slice; fill in the other ranks' slices (checkpointing)."""
        for t in (self.exp_avg, self.exp_avg_sq):
            own = torch.zeros_like(t)
            for lo, hi in self._regions:
                a, b = slice_of(lo, hi, self.rank, self.world)
                own[a:b] = t[a:b]
            dist.all_reduce(own)   # the owned slices are disjoint: the sum is their concatenation
            t.copy_(own)
