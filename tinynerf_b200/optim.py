"""Adam as one streaming pass per step.

`FusedAdam` is a torch.optim.Optimizer with torch.optim.Adam's hyper-parameters and state names
(`step`, `exp_avg`, `exp_avg_sq`), as configured by the reference at src/run.py:186 (L2 weight decay,
no amsgrad).  Every parameter tensor of every group goes through tnf_adam_step: p, g, m, v are read
once and p, m, v written once (28 B/param) instead of torch's ~10 foreach passes.
"""
from __future__ import annotations

import ctypes as C
import math

import torch

from . import _lib


class FusedAdam(torch.optim.Optimizer):
    def __init__(self, params, lr=1e-3, betas=(0.9, 0.999), eps=1e-8, weight_decay=0.0):
        super().__init__(params, dict(lr=lr, betas=betas, eps=eps, weight_decay=weight_decay))

    @torch.no_grad()
    def step(self, closure=None, subset=None, max_blocks: int = 0):
        """subset: optional collection of parameters -- only these are updated by this call (the trainer updates the
        feature planes on a second stream beside the heads' weight-gradient kernels and everything else afterwards);
        max_blocks > 0 caps the grid (persistent form, tnf_adam_step_grid)."""
        only = None if subset is None else frozenset(id(p) for p in subset)
        loss = None
        if closure is not None:
            with torch.enable_grad():
                loss = closure()
        _lib.load()
        tables = self.__dict__.setdefault("_tnf_tables", {})
        for group in self.param_groups:
            # fast path: same parameter / gradient / state tensor OBJECTS as in the previous call of this (group, subset) --
            # the fused iteration keeps every p.grad a view of one flat buffer, so nothing moves between steps and the ~100
            # data_ptr() calls + list building below are skipped
            fast = tables.get((id(group), only, "fast"))
            ok = fast is not None
            if ok:
                cnt = 0
                for p in group["params"]:
                    if p.grad is not None and (only is None or id(p) in only):
                        cnt += 1
                ok = cnt == fast["n"]
            if ok:
                for p, g, m, v, dp in zip(fast["ps"], fast["gs"], fast["ms"], fast["vs"], fast["dptr"]):
                    st = self.state[p]
                    if p.grad is not g or st["exp_avg"] is not m or st["exp_avg_sq"] is not v or p.data_ptr() != dp:
                        ok = False
                        break
            if ok:
                step = None
                for p in fast["ps"]:
                    st = self.state[p]
                    st["step"] += 1
                    step = st["step"] if step is None else step
                    if st["step"] != step:
                        raise RuntimeError("FusedAdam expects all parameters of a group to share the step count")
                b1, b2 = group["betas"]
                with torch.cuda.device(fast["ps"][0].device):
                    _lib.call("tnf_adam_step_grid", *fast["tabs"], fast["n"], float(group["lr"]), float(b1), float(b2),
                              float(group["eps"]), float(group["weight_decay"]), int(step), int(max_blocks), _lib.stream_ptr(),
                              nbytes=fast["nbytes"], extra_kernels=math.ceil(fast["n"] / 48) - 1)
                continue
            ps, gs, ms, vs = [], [], [], []
            step = None
            for p in group["params"]:
                if p.grad is None or (only is not None and id(p) not in only):
                    continue
                _lib.require_cuda(p, "parameter")
                if p.dtype != torch.float32:
                    raise RuntimeError("FusedAdam handles fp32 parameters only")
                st = self.state[p]
                if not st:
                    st["step"] = 0
                    st["exp_avg"] = torch.zeros_like(p, memory_format=torch.preserve_format)
                    st["exp_avg_sq"] = torch.zeros_like(p, memory_format=torch.preserve_format)
                st["step"] += 1
                step = st["step"] if step is None else step
                if st["step"] != step:
                    raise RuntimeError("FusedAdam expects all parameters of a group to share the step count")
                g = p.grad
                if g.stride() != p.stride():  # bring the gradient to the parameter's memory order
                    g = torch.empty_like(p, memory_format=torch.preserve_format).copy_(g)
                ps.append(p); gs.append(g); ms.append(st["exp_avg"]); vs.append(st["exp_avg_sq"])
            if not ps:
                continue
            n = len(ps)
            key = tuple(t.data_ptr() for ts in (ps, gs, ms, vs) for t in ts)
            cached = tables.get((id(group), only))
            if cached is None or cached[0] != key:  # pointer tables are rebuilt only when a tensor moved
                tab = lambda ts: (C.c_void_p * n)(*[t.data_ptr() for t in ts])
                cached = (key, tab(ps), tab(gs), tab(ms), tab(vs), (C.c_int64 * n)(*[t.numel() for t in ps]))
                tables[(id(group), only)] = cached
            _, tp, tg, tm, tv, numel = cached
            if all(g is p.grad for p, g in zip(ps, gs)):   # gradients already in the parameters' memory order: reusable as they are
                tables[(id(group), only, "fast")] = {"ps": list(ps), "gs": list(gs), "ms": list(ms), "vs": list(vs),
                                                     "dptr": [t.data_ptr() for t in ps], "tabs": (tp, tg, tm, tv, numel),
                                                     "n": n, "nbytes": 28 * sum(t.numel() for t in ps)}
            b1, b2 = group["betas"]
            with torch.cuda.device(ps[0].device):
                _lib.call("tnf_adam_step_grid", tp, tg, tm, tv, numel, n, float(group["lr"]), float(b1),
                          float(b2), float(group["eps"]), float(group["weight_decay"]), int(step), int(max_blocks), _lib.stream_ptr(),
                          nbytes=28 * sum(t.numel() for t in ps), extra_kernels=math.ceil(n / 48) - 1)
        return loss
