"""Diagnostics: wgrad_kernel MMA-warp wait/issue cycles (CTA 0)."""
import sys, ctypes as C
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
import torch
from tinynerf_b200 import _lib
lib = _lib.load()
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
buf = torch.zeros(32, dtype=torch.int64, device="cuda")
lib.tnf_debug_role_timing.argtypes = [C.c_void_p]
assert lib.tnf_debug_role_timing(buf.data_ptr()) == 0
m = 1 << 18
for (k, n) in ((64, 64), (96, 64), (148, 64)):
    ld = (k + 3) // 4 * 4
    x = torch.randn(m, ld, device="cuda"); dy = torch.randn(m, n, device="cuda")
    dw = torch.zeros(n, k, device="cuda"); db = torch.zeros(n, device="cuda")
    for rep in range(2):
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        flush.zero_()
        s.record()
        _lib.call("tnf_linear_bwd_weight", dy.data_ptr(), n, x.data_ptr(), ld, dw.data_ptr(), db.data_ptr(), m, n, k, _lib.stream_ptr())
        e.record(); torch.cuda.synchronize()
    v = buf.tolist()
    print(f"K={k}: {s.elapsed_time(e)*1e3:.1f} us | mma warp: wait_yfull {v[8]} wait_xfull {v[9]} issue {v[11]} total {v[10]} | producer: wait pempty {v[16]} xempty {v[17]} total {v[18]} | dY warps: wait pfull {v[19]} aempty {v[20]} work {v[21]} | X lo warps: wait land {v[22]} lempty {v[23]} work {v[25]} | CTA 0 clocks since entry: loop start {v[26]} first A {v[27]} loop end {v[28]} all MMAs done {v[29]} flushed {v[30]}")
