"""Diagnostics: a few launches of the weights op for `ncu -k regex:weights_` (sizes from argv, default 2^20 2^22)."""
import sys
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
import torch
from tinynerf_b200 import _cuda, synthetic

dev = torch.device("cuda")
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
for ln in [int(a) for a in sys.argv[1:]] or [20, 22]:
    n = 1 << ln
    sig, info, g = synthetic.packed_rays(n, seed=1000 + ln)
    sig, info, g = sig.to(dev), info.to(dev), g.to(dev)
    st = torch.full_like(sig, 5.196 / 256)
    for rep in range(2):
        flush.zero_()
        w = _cuda.weights_fwd(sig, st, info, 1e-4, _cuda.TRUSTED_PARTITION)
        flush.zero_()
        _cuda.weights_bwd(sig, st, info, w, g, _cuda.TRUSTED_PARTITION)
    torch.cuda.synchronize()
