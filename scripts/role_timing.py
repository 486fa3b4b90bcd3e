"""Diagnostics: where each warp role of linear_kernel waits (cycles, CTA 0)."""
import sys, ctypes as C
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
import torch
from tinynerf_b200 import _lib
lib = _lib.load()
dev = "cuda"
buf = torch.zeros(32, dtype=torch.int64, device=dev)
lib.tnf_debug_role_timing.argtypes = [C.c_void_p]
assert lib.tnf_debug_role_timing(buf.data_ptr()) == 0
m = 1 << 18
for (k, n) in ((64, 64), (148, 64), (96, 64)):
    x = torch.randn(m, k if k % 4 == 0 else k + (4 - k % 4), device=dev)
    w = torch.randn(n, k, device=dev) * 0.1
    b = torch.zeros(n, device=dev)
    y = torch.empty(m, n, device=dev)
    for rep in range(2):
        buf.zero_()
        buf[26] = 1 << 62
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record()
        _lib.call("tnf_linear_fwd", x.data_ptr(), x.stride(0), w.data_ptr(), b.data_ptr(), y.data_ptr(), n, m, n, k, 1, None, None, None, 0, 0, _lib.stream_ptr())
        e.record(); torch.cuda.synchronize()
    v = buf.tolist()
    print(f"K={k}: {s.elapsed_time(e)*1e3:.1f} us | loader: wait_hempty {v[0]} wait_cp {v[1]} wait_lempty {v[2]} lo_pass {v[3]} total {v[4]} items {v[5]}"
          f" entry->roles {v[6]} cyc | in-kernel span (min entry .. max exit over CTAs) {(v[25]-v[26])/1e3:.1f} us"
          f" | mma: wait_tempty {v[8]} wait_full {v[9]} total {v[10]} | epi: wait_tfull {v[16]} total {v[17]} tiles {v[18]}")

# back-to-back launches, no other kernel in between
x = torch.randn(m, 64, device=dev); w = torch.randn(64, 64, device=dev) * 0.1; b = torch.zeros(64, device=dev); y = torch.empty(m, 64, device=dev)
assert lib.tnf_debug_role_timing(None) == 0
for reps in (1, 4, 16):
    torch.cuda.synchronize()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    for _ in range(reps):
        _lib.call("tnf_linear_fwd", x.data_ptr(), 64, w.data_ptr(), b.data_ptr(), y.data_ptr(), 64, m, 64, 64, 1, None, None, None, 0, 0, _lib.stream_ptr())
    e.record(); torch.cuda.synchronize()
    print(f"{reps} back-to-back K=64 launches: {s.elapsed_time(e)*1e3/reps:.1f} us each")
