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
    assert len(nb.GROUPED_FUNCS) == 15 and len(nb.OTHER_FUNCS) == 2
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
