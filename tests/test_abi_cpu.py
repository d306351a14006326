"""CPU-only checks of the drop-in boundary: libnbg_b200.so loads without a GPU and exports
every symbol include/nbg_b200.h declares; the host mirror validates arguments like the
reference does (numbagg/decorators.py:305-341, 374-414, 465-487, 558-660) BEFORE touching
the device; and the product has no CPU fallback."""

import ctypes
import inspect
import os
import re

import numpy as np
import pytest

import numbagg_b200 as nb
from numbagg_b200 import _lib

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared_symbols():
    text = open(os.path.join(ROOT, "include", "nbg_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(nbg_[a-z0-9_]+)\s*\(", text)))


def test_library_exports_every_declared_symbol():
    assert os.path.exists(_lib.LIB_PATH), "run `python -m numbagg_b200.build` (or __graft_entry__.build())"
    handle = ctypes.CDLL(_lib.LIB_PATH)
    declared = _declared_symbols()
    assert len(declared) >= 14
    for name in declared:
        assert hasattr(handle, name), f"{name} declared in include/nbg_b200.h but not exported"
    assert set(declared) == set(_lib.EXPORTED_SYMBOLS)
    assert handle.nbg_abi_version() == 1


def test_no_cpu_fallback():
    import torch

    if torch.cuda.is_available():
        pytest.skip("GPU present")
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        nb.move_mean(np.arange(10.0), window=3)
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        nb.group_nansum(np.arange(4.0), np.array([0, 1, 0, 1]))


def test_product_never_imports_oracle():
    pkg = os.path.join(ROOT, "numbagg_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                text = open(os.path.join(dirpath, f)).read()
                assert "oracle" not in text, f"{f} mentions the oracle"


def test_function_census_and_metadata():
    assert len(nb.MOVE_FUNCS) == 6 and len(nb.MOVE_EXP_FUNCS) == 7
    assert len(nb.GROUPED_FUNCS) == 15 and len(nb.OTHER_FUNCS) == 2 and len(nb.AGGREGATION_FUNCS) == 11
    assert len(nb.MATRIX_FUNCS) == 6 and repr(nb.move_corrmatrix) == "numbagg.move_corrmatrix"
    assert len(nb.QUANTILE_FUNCS) == 2 and repr(nb.nanquantile) == "numbagg.nanquantile"
    assert repr(nb.nansum) == "numbagg.nansum" and nb.nanvar.supports_ddof and not nb.nansum.supports_ddof
    assert repr(nb.move_mean) == "numbagg.move_mean"  # numbagg/decorators.py:119-120
    assert nb.group_nanvar.supports_ddof and not nb.group_nanvar.supports_ints and not nb.group_nanvar.supports_bool
    assert nb.group_nanmean.supports_bool and not nb.group_nanmean.supports_ints
    sig = inspect.signature(nb.move_mean)
    assert sig.parameters["window"].kind is inspect.Parameter.KEYWORD_ONLY
    assert sig.parameters["min_count"].default is None and sig.parameters["axis"].default == -1
    sig = inspect.signature(nb.move_exp_nanmean)
    assert sig.parameters["alpha"].kind is inspect.Parameter.KEYWORD_ONLY and sig.parameters["min_weight"].default == 0
    sig = inspect.signature(nb.ffill)
    assert sig.parameters["limit"].default is None
    sig = inspect.signature(nb.group_nansum)
    assert sig.parameters["ddof"].default == 1 and sig.parameters["num_labels"].default is None


def test_every_public_function_of_the_reference_exists():
    """numbagg.__all__ (numbagg/__init__.py:3-62), frozen here because the reference cannot be
    imported on the GPU box."""
    reference_all = """allnan anynan bfill count ffill group_nanall group_nanany group_nanargmax group_nanargmin
    group_nancount group_nanfirst group_nanlast group_nanmax group_nanmean group_nanmin group_nanprod group_nanstd
    group_nansum group_nansum_of_squares group_nanvar move_corr move_corrmatrix move_cov move_covmatrix
    move_exp_nancorr move_exp_nancorrmatrix move_exp_nancount move_exp_nancov move_exp_nancovmatrix move_exp_nanmean
    move_exp_nanstd move_exp_nansum move_exp_nanvar move_mean move_std move_sum move_var nanargmax nanargmin
    nancorrmatrix nancount nancovmatrix nanmax nanmean nanmedian nanmin nanquantile nanstd nansum nanvar""".split()
    missing = [name for name in reference_all if not callable(getattr(nb, name, None))]
    assert not missing, missing
    assert nb.count is nb.nancount


A = np.arange(10.0)


def test_move_validation():
    # test_moving.py:168-177 and decorators.py:313-334
    with pytest.raises(TypeError):
        nb.move_mean(A, 3)  # window is keyword-only
    with pytest.raises(ValueError, match="window not in valid range"):
        nb.move_mean(A, window=0)
    with pytest.raises(ValueError, match="window not in valid range"):
        nb.move_mean(A, window=11)
    with pytest.raises(ValueError, match="min_count must be positive"):
        nb.move_mean(A, window=3, min_count=-1)
    with pytest.raises(TypeError):
        nb.move_mean(A, window=2.5)
    with pytest.raises(ValueError, match="only one axis"):
        nb.move_mean(np.zeros((3, 4)), window=2, axis=(0, 1))
    with pytest.raises(ValueError, match="empty tuple"):
        nb.move_cov(A, A, window=2, axis=())
    assert nb.move_mean(A, window=3, axis=()) is A  # returns the input object itself


def test_move_exp_validation():
    with pytest.raises(TypeError):
        nb.move_exp_nanmean(A, halflife=3)  # only alpha exists (SURVEY discrepancy table)
    with pytest.raises(ValueError, match="Only one axis"):
        nb.move_exp_nanmean(np.zeros((3, 4)), alpha=0.5, axis=(0, 1))
    assert nb.move_exp_nansum(A, alpha=0.5, axis=()) is A


def test_fill_validation():
    with pytest.raises(ValueError, match="`limit` must be positive"):
        nb.ffill(A, limit=-1)
    with pytest.raises(TypeError, match="Unsupported dtype for fill operation"):
        nb.ffill(np.array(["a", "b"]))
    # integers have no NaN: identity, no device needed (funcs.py:303,319)
    ints = np.arange(5, dtype=np.int32)
    out = nb.bfill(ints, limit=1)
    assert out.dtype == np.int32 and np.array_equal(out, ints) and out is not ints


def test_group_validation():
    v = np.arange(12.0).reshape(4, 3)
    with pytest.raises(TypeError, match="labels must be an integer array"):
        nb.group_nansum(v[0], np.array([0.0, 1.0, 0.0]))
    with pytest.raises(TypeError, match="labels must be an integer array"):
        nb.group_nansum(v[0], np.array([0, 1, 0], dtype=np.uint8))
    with pytest.raises(ValueError, match="axis required"):
        nb.group_nansum(v, np.array([0, 1, 0]))
    with pytest.raises(ValueError, match="must have same shape along axis"):
        nb.group_nansum(v, np.array([0, 1, 0]), axis=0)
    with pytest.raises(ValueError, match="must have same shape along axis"):
        nb.group_nansum(v, np.zeros((3, 3), dtype=np.int64), axis=(0, 1))
    with pytest.raises(TypeError, match="does not support boolean input"):
        nb.group_nanvar(np.array([True, False]), np.array([0, 0]))


def test_c_abi_rejects_bad_arguments_without_a_device():
    """Argument errors are reported through return codes + nbg_last_error() before any CUDA
    call is made (numbagg raises Python exceptions before dispatch; the C layer never throws)."""
    L = _lib.lib()
    err = lambda: L.nbg_last_error().decode()
    dummy = ctypes.c_void_p(256)
    # window must be positive; dtype must be a float type; negative sizes
    assert L.nbg_move(0, _lib.NBG_F64, dummy, None, dummy, 1, 10, 1, 0, 1, None, None, 0, None) == -3 and "window" in err()
    assert L.nbg_move(0, _lib.NBG_I32, dummy, None, dummy, 1, 10, 1, 3, 1, None, None, 0, None) == -1 and "dtype" in err()
    assert L.nbg_move(0, _lib.NBG_F64, dummy, None, dummy, -1, 10, 1, 3, 1, None, None, 0, None) == -3
    assert L.nbg_move(4, _lib.NBG_F64, dummy, None, dummy, 1, 10, 1, 3, 1, None, None, 0, None) == -3  # cov needs b
    assert L.nbg_move(99, _lib.NBG_F64, dummy, None, dummy, 1, 10, 1, 3, 1, None, None, 0, None) == -2
    assert L.nbg_move(0, _lib.NBG_F64, dummy, None, dummy, 1, 10, 1, 3, 1, None, None, 5, None) == -3  # halo_len without halo
    # empty problems are fine and launch nothing
    before = L.nbg_launch_count()
    assert L.nbg_move(0, _lib.NBG_F64, None, None, None, 0, 10, 1, 3, 1, None, None, 0, None) == 0
    assert L.nbg_launch_count() == before
    assert L.nbg_fill(7, 8, dummy, dummy, 1, 10, 1, 3, None, None, None, 0, None) == -2
    assert L.nbg_fill(0, 2, dummy, dummy, 1, 10, 1, 3, None, None, None, 0, None) == -1 and "itemsize" in err()
    assert L.nbg_fill(0, 8, dummy, dummy, 1, 10, 1, -1, None, None, None, 0, None) == -3 and "limit" in err()
    assert L.nbg_fill(0, 8, dummy, None, 1, 10, 1, 3, None, None, None, 0, None) == -3  # neither out nor agg_out
    assert L.nbg_move_exp(0, _lib.NBG_I64, dummy, None, None, 0, 0.5, 0.0, dummy, 1, 10, 1, None, None, None, 0, None) == -1
    assert L.nbg_move_exp(5, _lib.NBG_F64, dummy, None, None, 0, 0.5, 0.0, dummy, 1, 10, 1, None, None, None, 0, None) == -3  # cov needs a2
    # grouped: float-only ops refuse integer values; workspace size is checked
    assert L.nbg_group_accumulate(_lib.GROUP_OPS["group_nanvar"], _lib.NBG_I32, _lib.NBG_I64, dummy, dummy, 0, dummy,
                                  1 << 30, 1, 10, 4, 0, None) == -1
    assert L.nbg_group(_lib.GROUP_OPS["group_nansum"], _lib.NBG_F64, _lib.NBG_I64, dummy, dummy, 0, dummy, 1, 10, 4, 1,
                       dummy, 8, None) == -6 and "workspace" in err()
    assert L.nbg_group_record_words(_lib.GROUP_OPS["group_nanvar"]) == 4
    assert L.nbg_group_record_words(_lib.GROUP_OPS["group_nansum"]) == 1
    assert L.nbg_group_workspace_bytes(1, _lib.NBG_F32, 10, 1000, 7) >= 10 * 7 * 8
    # plain reductions
    R = _lib.REDUCE_OPS
    assert L.nbg_reduce(99, _lib.NBG_F64, dummy, dummy, 1, 10, 1, 1, None, 0, None) == -2
    assert L.nbg_reduce(R["nansum"], 7, dummy, dummy, 1, 10, 1, 1, None, 0, None) == -1
    assert L.nbg_reduce(R["nanmean"], _lib.NBG_I64, dummy, dummy, 1, 10, 1, 1, None, 0, None) == -1 and "float" in err()
    assert L.nbg_reduce(R["nansum"], _lib.NBG_F64, dummy, None, 1, 10, 1, 1, None, 0, None) == -3
    assert L.nbg_reduce(R["nansum"], _lib.NBG_F64, dummy, dummy, -1, 10, 1, 1, None, 0, None) == -3
    # one long row is cut into segments: the partial states need the advertised workspace
    need = L.nbg_reduce_workspace_bytes(R["nansum"], _lib.NBG_F64, 1, 10_000_000, 1)
    assert need >= 3 * 8 * 2
    assert L.nbg_reduce(R["nansum"], _lib.NBG_F64, dummy, dummy, 1, 10_000_000, 1, 1, dummy, 8, None) == -6 and "workspace" in err()
    assert L.nbg_reduce_workspace_bytes(R["nansum"], _lib.NBG_F64, 100_000, 100, 1) == 0  # one CTA pass, no partials
    # quantiles
    assert L.nbg_quantile(dummy, dummy, dummy, 3, 10, 17, None, 0, None) == -3 and "16 quantiles" in err()
    assert L.nbg_quantile(dummy, None, dummy, 3, 10, 2, None, 0, None) == -3
    assert L.nbg_quantile(dummy, dummy, dummy, 3, 100_000, 2, None, 0, None) == -6 and "workspace" in err()
    assert L.nbg_quantile_workspace_bytes(3, 4096, 2) == 0 and L.nbg_quantile_workspace_bytes(3, 4097, 2) > 3 * 4 * 1024
    # matrix functions
    M = _lib.MATRIX_OPS
    assert L.nbg_matrix(9, _lib.NBG_F64, dummy, None, 0, 0.0, dummy, 1, 10, 3, 2, 1, None) == -2
    assert L.nbg_matrix(M["nancovmatrix"], _lib.NBG_I64, dummy, None, 0, 0.0, dummy, 1, 10, 3, 0, 0, None) == -1
    assert L.nbg_matrix(M["move_covmatrix"], _lib.NBG_F64, dummy, None, 0, 0.0, dummy, 1, 10, 3, 0, 1, None) == -3 and "window" in err()
    assert L.nbg_matrix(M["move_exp_nancovmatrix"], _lib.NBG_F64, dummy, None, 0, 0.0, dummy, 1, 10, 3, 0, 0, None) == -3 and "alpha" in err()
    assert L.nbg_matrix(M["nancorrmatrix"], _lib.NBG_F32, dummy, None, 0, 0.0, None, 1, 10, 3, 0, 0, None) == -3
    before = L.nbg_launch_count()
    assert L.nbg_matrix(M["nancorrmatrix"], _lib.NBG_F32, None, None, 0, 0.0, None, 0, 10, 3, 0, 0, None) == 0
    assert L.nbg_quantile(None, None, None, 0, 10, 2, None, 0, None) == 0
    assert L.nbg_reduce(R["nansum"], _lib.NBG_F64, None, None, 0, 10, 1, 1, None, 0, None) == 0  # no outputs
    assert L.nbg_reduce_merge(R["nanvar"], _lib.NBG_F64, None, 0, 0, None, 0, 1, None) == 0
    assert L.nbg_launch_count() == before


def test_c_abi_header_constants_match_python_binding():
    text = open(os.path.join(ROOT, "include", "nbg_b200.h")).read()
    for name, value in (("NBG_EXP_STATE", _lib.NBG_EXP_STATE), ("NBG_FILL_STATE", _lib.NBG_FILL_STATE),
                        ("NBG_GROUP_WS_CHANNELS", _lib.NBG_GROUP_WS_CHANNELS), ("NBG_ABI_VERSION", 1),
                        ("NBG_REDUCE_STATE_WORDS", _lib.NBG_REDUCE_STATE_WORDS)):
        m = re.search(rf"#define\s+{name}\s+(\d+)", text)
        assert m and int(m.group(1)) == value, name
    for table, prefix in ((_lib.MOVE_OPS, "NBG_"), (_lib.EXP_OPS, "NBG_"), (_lib.GROUP_OPS, "NBG_"),
                          (_lib.REDUCE_OPS, "NBG_RED_"), (_lib.MATRIX_OPS, "NBG_MAT_")):
        for fname, code in table.items():
            enum = prefix + fname.upper().replace("MOVE_EXP_", "EXP_")
            if prefix == "NBG_MAT_":  # NBG_MAT_NANCORR, NBG_MAT_MOVE_COV, NBG_MAT_EXP_CORR ...
                enum = prefix + fname.upper().replace("MATRIX", "").replace("MOVE_EXP_NAN", "EXP_")
            m = re.search(rf"\b{enum}\s*=\s*(\d+)", text)
            assert m and int(m.group(1)) == code, enum
