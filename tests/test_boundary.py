"""The drop-in boundary (CPU side): the C-ABI library loads and exports every symbol the header
declares, the product never touches oracle/, and the Python shim raises the reference's errors."""
import ctypes
import re
from pathlib import Path

import pytest
import torch

ROOT = Path(__file__).resolve().parents[1]


def header_symbols():
    text = (ROOT / "include" / "tinynerf_b200.h").read_text()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(tnf_[a-z0-9_]+)\s*\(", text)))


def test_library_exports_every_declared_symbol():
    import __graft_entry__
    __graft_entry__.build()
    from tinynerf_b200 import _lib
    lib = ctypes.CDLL(str(_lib.lib_path()))
    syms = header_symbols()
    assert len(syms) >= 15
    for s in syms:
        assert hasattr(lib, s), f"{s} declared in include/tinynerf_b200.h but not exported"
    assert sorted(_lib.declared_symbols()) == syms, "ctypes signatures out of sync with the header"
    assert _lib.load().tnf_version() >= 1000


def test_product_never_imports_the_oracle():
    for p in (ROOT / "tinynerf_b200").rglob("*"):
        if p.suffix in {".py", ".cu", ".cuh", ".h"}:
            text = p.read_text()
            assert "oracle" not in text.replace("# oracle", ""), f"{p} mentions the oracle"


def test_shim_rejects_cpu_tensors_like_the_reference():
    from tinynerf_b200 import _cuda
    s = torch.ones(4)
    i = torch.zeros(1, 2, dtype=torch.int32)
    with pytest.raises(RuntimeError, match="sigmas must be a CUDA tensor"):   # src/cuda.cu:62
        _cuda.compute_weights_fwd(s, s, i, 1e-4)
    with pytest.raises(RuntimeError, match="sigmas must be a CUDA tensor"):
        _cuda.compute_weights_bwd(s, s, i, s, s)


def test_no_cpu_fallback_in_core():
    from tinynerf_b200 import core, models
    grid = core.OccupancyGrid(8, 0.1)
    with pytest.raises(RuntimeError, match="CUDA"):
        grid(torch.zeros(4, 3))
    field = models.KPlanesFeatureField(32)
    with pytest.raises(RuntimeError, match="CUDA"):
        field(torch.zeros(4, 3))


def test_c_abi_argument_validation_without_gpu():
    """Pure argument checks return TNF_E_INVALID before any CUDA call."""
    from tinynerf_b200 import _lib
    lib = _lib.load()
    assert lib.tnf_weights_fwd(None, None, 1, None, 0.0, None, -1, 0, 0, None, None) == -1
    assert b"negative" in lib.tnf_last_error()
    assert lib.tnf_weights_fwd(None, None, 1, None, 0.0, None, 0, 0, 0, None, None) == 0  # empty: nothing to do
    assert lib.tnf_kplanes_fwd(None, None, 0, 32, None, 3, 0, None, None) == -1
