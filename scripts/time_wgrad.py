"""Diagnostics: tnf_linear_bwd_weight on the five head-layer shapes of a K-Planes step (M = 2^18), stacked-SS kernel
(TNF_WGRAD=ss) against the tensor-memory-A / TMA kernel (default)."""
import os, sys
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
import torch
from tinynerf_b200 import _lib

m = 1 << 18
shapes = ((64, 64), (148, 64), (96, 64))
bufs = {}
for k, n in shapes:
    ld = (k + 3) // 4 * 4
    bufs[(k, n)] = (torch.randn(m, ld, device="cuda"), torch.randn(m, n, device="cuda"), torch.zeros(n, k, device="cuda"),
                    torch.zeros(n, device="cuda"), ld)
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
for mode in (sys.argv[1:] or ["ss", "tma", "ss", "tma"]):
    _lib.load().tnf_set_variant(0, 1 if mode == "ss" else 0)
    for k, n in shapes:
        x, dy, dw, db, ld = bufs[(k, n)]
        ts = []
        for rep in range(6):
            flush.zero_()
            s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            s.record()
            _lib.call("tnf_linear_bwd_weight", dy.data_ptr(), n, x.data_ptr(), ld, dw.data_ptr(), db.data_ptr(), m, n, k, _lib.stream_ptr())
            e.record(); torch.cuda.synchronize()
            ts.append(s.elapsed_time(e) * 1e3)
        ts = sorted(ts[1:])
        gb = 4 * m * (n + k) / 1e9
        print(f"{mode} K={k:3d}: median {ts[len(ts)//2]:7.1f} us  min {ts[0]:7.1f} us  -> {gb / (ts[len(ts)//2] * 1e-6):7.0f} GB/s", flush=True)
