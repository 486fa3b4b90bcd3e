"""Diagnostics: a 128 x 128 layer (Cobafa trunk) at M = 2^18 rows, forward and data gradient: the weight-stationary kernel
(csrc/wstat.cu, default) against linear_kernel with the weights resident in shared memory, and the one-pass 128-wide weight
gradient (wgrad128_tma_kernel) against two half-launches of wgrad_tma_kernel (tnf_set_variant(3, 1) selects the older kernels)."""
import sys
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
import torch
from tinynerf_b200 import _lib

m = int(sys.argv[1]) if len(sys.argv) > 1 else 1 << 18
dev = "cuda"
x = torch.randn(m, 128, device=dev).relu()
w = torch.randn(128, 128, device=dev) / 128 ** 0.5
b = torch.randn(128, device=dev)
y = torch.empty(m, 128, device=dev)
dx = torch.empty(m, 128, device=dev)
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
lib = _lib.load()


def timed(fn):
    ts = []
    for _ in range(7):
        flush.zero_()
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record(); fn(); e.record()
        torch.cuda.synchronize()
        ts.append(s.elapsed_time(e) * 1e3)
    ts = sorted(ts[1:])
    return ts[len(ts) // 2], ts[0]


for variant, label in ((1, "resident (linear_kernel)"), (0, "weight-stationary (wstat)"), (1, "resident (linear_kernel)"), (0, "weight-stationary (wstat)")):
    lib.tnf_set_variant(3, variant)
    f = timed(lambda: _lib.call("tnf_linear_fwd", x.data_ptr(), 128, w.data_ptr(), b.data_ptr(), y.data_ptr(), 128, m, 128, 128, 1,
                                None, None, None, 0, 0, _lib.stream_ptr()))
    d = timed(lambda: _lib.call("tnf_linear_bwd_data", y.data_ptr(), 128, w.data_ptr(), dx.data_ptr(), 128, x.data_ptr(), 128, m, 128, 128,
                                _lib.stream_ptr()))
    gw, gbias = torch.zeros(128, 128, device=dev), torch.zeros(128, device=dev)
    w_ = timed(lambda: _lib.call("tnf_linear_bwd_weight", y.data_ptr(), 128, x.data_ptr(), 128, gw.data_ptr(), gbias.data_ptr(), m, 128, 128,
                                 _lib.stream_ptr()))
    gb_f, gb_d = 4 * m * 256 / 1e9, 4 * m * 384 / 1e9
    print(f"{label:28s} fwd median {f[0]:7.1f} us (min {f[1]:7.1f}) {gb_f / f[0] * 1e6:6.0f} GB/s | "
          f"dgrad median {d[0]:7.1f} us (min {d[1]:7.1f}) {gb_d / d[0] * 1e6:6.0f} GB/s | "
          f"wgrad median {w_[0]:7.1f} us (min {w_[1]:7.1f}) {gb_f / w_[0] * 1e6:6.0f} GB/s", flush=True)
lib.tnf_set_variant(3, 0)
