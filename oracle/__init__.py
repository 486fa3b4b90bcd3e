"""oracle -- CPU restatement of the reference algorithm.  TEST INFRASTRUCTURE ONLY.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may import
this package; nothing under tinynerf_b200/ does (tests/test_boundary.py enforces it).

    oracle.build()          gcc -> oracle/_build/libtnf_oracle.so   (plus oracle/_ref when /root/reference exists)
    oracle.c                ctypes wrappers over tnf_oracle.c working on CPU torch tensors
    oracle.ref_port         functional PyTorch restatement of src/core.py + src/models.py (any device)
    oracle.load_ref_cuda()  the UNMODIFIED reference `_cuda` op from oracle/_ref/_cuda.so (GPU only)
"""
from __future__ import annotations

import importlib.util
import subprocess
from pathlib import Path

HERE = Path(__file__).resolve().parent
SRC = HERE / "tnf_oracle.c"
OUT = HERE / "_build" / "libtnf_oracle.so"
REF_SO = HERE / "_ref" / "_cuda.so"


def build(force: bool = False) -> Path:
    OUT.parent.mkdir(exist_ok=True)
    if force or not OUT.exists() or OUT.stat().st_mtime < SRC.stat().st_mtime:
        cmd = ["gcc", "-O2", "-ffp-contract=off", "-fPIC", "-shared", "-o", str(OUT), str(SRC), "-lm"]
        subprocess.run(cmd, check=True)
    return OUT


def load_ref_cuda():
    """Import the reference's own pybind module built by oracle/build_ref.py, or None if absent."""
    if not REF_SO.exists():
        return None
    import torch  # noqa: F401  (the extension links against libtorch)
    spec = importlib.util.spec_from_file_location("_cuda", str(REF_SO))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod
