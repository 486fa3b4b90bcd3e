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
    # entry points added for the training iteration: shape / layout errors are reported before any CUDA call
    one = ctypes.c_void_p(16)   # a non-null, 16-byte aligned dummy pointer that is never dereferenced
    assert lib.tnf_linear_bwd_weight_cat(one, 64, one, 52, 51, one, 96, 96, one, one, 1000, 128, None, None) == -1
    assert b"64 output features" in lib.tnf_last_error()
    assert lib.tnf_linear_bwd_weight_cat(one, 64, one, 52, 51, one, 96, 96, one, one, 0, 64, None, None) == 0   # m == 0
    tab = (ctypes.c_void_p * 5)(16, 16, 16, 16, 16)
    assert lib.tnf_heads_fwd(one, 96, 96, one, 52, 147, 50, tab, tab, tab, tab, None, None, one, one, 10, one, None) == -1
    assert b"xc_cols" in lib.tnf_last_error()
    assert lib.tnf_composite_loss_fwd_bwd(one, one, one, 10, 0, None, one, 1.0, None, 1.0, one, one, one, one, one, None, None, 0,
                                          None) == -1
    # several weight gradients in one launch: the job tables are checked on the host
    vp, i64, i32 = (ctypes.c_void_p * 2)(16, 16), (ctypes.c_int64 * 2)(64, 64), (ctypes.c_int32 * 2)(64, 96)
    assert lib.tnf_linear_bwd_weight_multi(0, vp, i64, vp, i64, i32, vp, None, 1000, None) == -1
    assert b"n_jobs" in lib.tnf_last_error()
    assert lib.tnf_linear_bwd_weight_multi(5, vp, i64, vp, i64, i32, vp, None, 1000, None) == -1
    assert lib.tnf_linear_bwd_weight_multi(2, vp, i64, vp, i64, i32, vp, None, 0, None) == 0        # m == 0: nothing to do
    assert lib.tnf_linear_bwd_weight_multi(2, vp, i64, vp, i64, (ctypes.c_int32 * 2)(64, 160), vp, None, 1000, None) == -1
    assert b"in_features" in lib.tnf_last_error()
    assert lib.tnf_linear_bwd_weight_multi(2, vp, i64, vp, (ctypes.c_int64 * 2)(64, 94), i32, vp, None, 1000, None) == -1
    assert b"leading dimensions" in lib.tnf_last_error()
    assert lib.tnf_set_variant(3, 1) == 0 and lib.tnf_set_variant(3, 0) == 1   # 128-wide layers: kernel-generation switch
    assert lib.tnf_wgrad_cat_scratch_bytes(51, 96) == (64 * 32 * 5 + 4) * 4
    assert lib.tnf_heads_workspace_bytes(96, 147) >= (6 + 2 + 3 + 3) * 16384


def test_host_helpers_of_the_abi():
    """tnf_shuffle_next (lazily shuffled ray order) and tnf_set_sm_budget run on the host."""
    from tinynerf_b200 import _lib
    lib = _lib.load()
    n, world = 1000, 4
    outs = []
    for rank in range(world):   # every rank walks the same seeded sequence and keeps the positions g % world == rank
        perm = torch.arange(n, dtype=torch.int64)
        fresh, rng = ctypes.c_int64(0), ctypes.c_uint64(12345)
        out = torch.empty(n // world, dtype=torch.int64)
        assert lib.tnf_shuffle_next(perm.data_ptr(), n, 0, n // world, rank, world, ctypes.byref(fresh), ctypes.byref(rng),
                                    out.data_ptr()) == 0
        assert fresh.value == n and sorted(perm.tolist()) == list(range(n))
        outs.append(out)
    allv = torch.cat(outs)
    assert sorted(allv.tolist()) == list(range(n)) and allv.tolist() != list(range(n))
    # positions below fresh_from are replayed without consuming random numbers
    perm = torch.arange(n, dtype=torch.int64)
    fresh, rng = ctypes.c_int64(0), ctypes.c_uint64(7)
    a, b = torch.empty(300, dtype=torch.int64), torch.empty(100, dtype=torch.int64)
    lib.tnf_shuffle_next(perm.data_ptr(), n, 0, 300, 0, 1, ctypes.byref(fresh), ctypes.byref(rng), a.data_ptr())
    state = rng.value
    lib.tnf_shuffle_next(perm.data_ptr(), n, 200, 100, 0, 1, ctypes.byref(fresh), ctypes.byref(rng), b.data_ptr())
    assert torch.equal(b, a[200:]) and rng.value == state and fresh.value == 300
    assert lib.tnf_shuffle_next(None, n, 0, 1, 0, 1, ctypes.byref(fresh), ctypes.byref(rng), b.data_ptr()) == -1
    prev = lib.tnf_set_sm_budget(100)
    assert lib.tnf_set_sm_budget(prev) == 100
