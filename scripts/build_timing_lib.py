"""Diagnostics: build scripts/ubench/libtnf_timing.so = the library with -DTNF_ROLE_TIMING (per-role cycle counters in the
tensor-core kernels).  Scripts select it with TNF_LIB_PATH."""
import subprocess, sys
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
from tinynerf_b200 import build as tb

out = Path(__file__).resolve().parent / "ubench"
obj = out / "timing_obj"
obj.mkdir(parents=True, exist_ok=True)
objs = []
for s in sorted(tb.CSRC.glob("*.cu")):
    o = obj / (s.stem + ".o")
    subprocess.run([tb._nvcc(), *tb.NVCC_FLAGS, "-DTNF_ROLE_TIMING", "-c", str(s), "-o", str(o)], check=True)
    objs.append(str(o))
subprocess.run([tb._nvcc(), "-shared", "-gencode", "arch=compute_100a,code=sm_100a", "-o", str(out / "libtnf_timing.so"), *objs], check=True)
print(out / "libtnf_timing.so")
