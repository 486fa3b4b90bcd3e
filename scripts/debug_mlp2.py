import sys
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
import torch
from tinynerf_b200 import _lib, mlp_ops
DEV = "cuda"
torch.manual_seed(0)
def rel(a, b): return ((a.double() - b.double()).abs().max() / b.double().abs().max()).item()
for m in (1000, 5000, 5120, 40000):
    g = torch.Generator().manual_seed(m)
    k, n = 96, 64
    x = (torch.randn(m, k, generator=g) * 0.5).to(DEV)
    w0 = (torch.randn(n, k, generator=g) / k ** 0.5).to(DEV); b0 = torch.randn(n, generator=g).to(DEV) * 0.1
    wh = (torch.randn(1, n, generator=g) / n ** 0.5).to(DEV); bh = torch.zeros(1, device=DEV)
    h, out = mlp_ops._lin_fwd(x, w0, b0, True, head=(wh, bh), head_act=1)
    h_ref = torch.relu(x.double() @ w0.double().t() + b0.double())
    out_ref = torch.exp(h_ref @ wh.double().t() + bh.double() - 1)
    go = torch.randn(m, 1, generator=g).to(DEV)
    dh = torch.empty_like(h); gwh = torch.zeros_like(wh); gbh = torch.zeros(1, device=DEV)
    with torch.cuda.device(0):
        _lib.call("tnf_head_bwd", h.data_ptr(), n, wh.data_ptr(), out.data_ptr(), go.data_ptr(), dh.data_ptr(), gwh.data_ptr(), gbh.data_ptr(), m, n, 1, 1, _lib.stream_ptr())
    dpre = go.double() * out_ref
    dh_ref = (dpre @ wh.double()) * (h_ref > 0)
    bad_rows = ((dh.double() - dh_ref).abs().max(1).values > 1e-5 * dh_ref.abs().max()).nonzero().flatten()
    print(m, "fwd h", rel(h, h_ref), "out", rel(out, out_ref), "dh", rel(dh, dh_ref), "bad rows", bad_rows[:10].tolist(), len(bad_rows))
    gw = torch.zeros_like(w0); gb = torch.zeros(n, device=DEV)
    dx = torch.empty(m, k, device=DEV)
    with torch.cuda.device(0):
        _lib.call("tnf_linear_bwd_weight", dh.data_ptr(), n, x.data_ptr(), k, gw.data_ptr(), gb.data_ptr(), m, n, k, _lib.stream_ptr())
        _lib.call("tnf_linear_bwd_data", dh.data_ptr(), n, w0.data_ptr(), dx.data_ptr(), k, None, 0, m, n, k, _lib.stream_ptr())
    gw_ref = dh.double().t() @ x.double(); gb_ref = dh.double().sum(0); dx_ref = dh.double() @ w0.double()
    bad = ((dx.double() - dx_ref).abs().max(1).values > 1e-5 * dx_ref.abs().max()).nonzero().flatten()
    print("   wgrad", rel(gw, gw_ref), "db", rel(gb, gb_ref), "dgrad", rel(dx, dx_ref), "bad dx rows", bad[:12].tolist(), len(bad))
    # repeat to see determinism
    gw2 = torch.zeros_like(w0); gb2 = torch.zeros(n, device=DEV)
    with torch.cuda.device(0):
        _lib.call("tnf_linear_bwd_weight", dh.data_ptr(), n, x.data_ptr(), k, gw2.data_ptr(), gb2.data_ptr(), m, n, k, _lib.stream_ptr())
    print("   wgrad again", rel(gw2, gw_ref), "equal", torch.equal(gw, gw2))
