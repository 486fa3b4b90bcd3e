import inspect, sys, re
sys.path.insert(0, '/root/repo')
import torch
from tinynerf_b200 import _lib
import tests.test_gpu_mlp as T
src = inspect.getsource(T.test_wide_stacks_forward_backward_vs_float64)
src = src[src.index("def "):]
src = src.replace("if e > 2e-5:", "if e > 0:")
ns = dict(T.__dict__)
exec(src, ns)
fn = ns["test_wide_stacks_forward_backward_vs_float64"]
for variant in (1, 0):
    _lib.load().tnf_set_variant(3, variant)
    for m in (3000,):
        try:
            fn(m)
        except AssertionError as ex:
            msg = str(ex)
            d = eval(msg[msg.index("{"):msg.index("}") + 1])
            top = sorted(d.items(), key=lambda kv: -kv[1])[:6]
            print("variant", variant, "m", m, [(k, f"{v:.2e}") for k, v in top], flush=True)
_lib.load().tnf_set_variant(3, 0)
