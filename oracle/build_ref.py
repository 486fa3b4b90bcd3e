"""Build recipe for oracle/_ref: the UNMODIFIED reference weights op compiled for sm_100a.

TEST INFRASTRUCTURE ONLY (see oracle/README.md). Compiles /root/reference/src/cuda.cu where it lies
(nothing is copied into this repo) with torch's cpp_extension, exactly like the reference's own
`load(name="_cuda", sources=['src/cuda.cu'])` (reference src/core.py:7), but with the build directory
pointed at oracle/_ref/ so the resulting `_cuda.so` travels to the GPU box with the snapshot.

Usage (only in the build container, where /root/reference exists):
    python oracle/build_ref.py
"""
import os
import sys
from pathlib import Path

REF = Path(os.environ.get("TNF_REFERENCE_ROOT", "/root/reference"))
OUT = Path(__file__).resolve().parent / "_ref"


def build(verbose: bool = False) -> Path | None:
    src = REF / "src" / "cuda.cu"
    if not src.exists():
        return None
    OUT.mkdir(exist_ok=True)
    so = OUT / "_cuda.so"
    if so.exists() and so.stat().st_mtime >= src.stat().st_mtime:
        return so
    os.environ.setdefault("TORCH_CUDA_ARCH_LIST", "10.0a")
    from torch.utils.cpp_extension import load
    load(name="_cuda", sources=[str(src)], build_directory=str(OUT), verbose=verbose)
    return so if so.exists() else None


if __name__ == "__main__":
    p = build(verbose=True)
    print("oracle/_ref:", p)
    sys.exit(0 if p else 1)
