"""numpy front-end of the CPU oracle (TEST INFRASTRUCTURE ONLY -- see nbg_oracle.c header).

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s CPU-baseline legs may import
this module.  ``numbagg_b200`` never does: the product has no CPU path.

The functions mirror the call signatures of the reference's public API
(numbagg/__init__.py:3-62) closely enough for parity tests to read like the reference's own
tests, but they do no argument validation: that is the product's job
(numbagg_b200/decorators.py) and is tested there.
"""

from __future__ import annotations

import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "_build", "libnbg_oracle.so")


def build(force: bool = False) -> str:
    """Compile nbg_oracle.c with the committed Makefile (gcc, -ffp-contract=off)."""
    src = os.path.join(_HERE, "nbg_oracle.c")
    if force or not os.path.exists(_SO) or os.path.getmtime(_SO) < os.path.getmtime(src):
        subprocess.run(["make", "-C", _HERE], check=True, capture_output=True)
    return _SO


_lib = None


def lib() -> ctypes.CDLL:
    global _lib
    if _lib is None:
        # libgomp's default active spin-wait collapses on shared/virtualised cores (60 ms
        # instead of 1.5 ms per call in the dev container): park idle workers instead.
        os.environ.setdefault("OMP_WAIT_POLICY", "passive")
        os.environ.setdefault("GOMP_SPINCOUNT", "0")
        _lib = ctypes.CDLL(build())
    return _lib


_SUFFIX = {
    np.dtype(np.float32): "f32",
    np.dtype(np.float64): "f64",
    np.dtype(np.int32): "i32",
    np.dtype(np.int64): "i64",
}

_i64 = ctypes.c_int64
_vp = ctypes.c_void_p


def _ptr(a: np.ndarray) -> ctypes.c_void_p:
    return ctypes.c_void_p(a.ctypes.data)


def _float_loop_dtype(*arrs) -> np.dtype:
    """NumPy's gufunc loop selection for the (float32, float64) loops the reference
    registers: float32 stays, float16 -> float32, everything else -> float64."""
    dt = np.result_type(*[np.asarray(a).dtype for a in arrs])
    if dt == np.float32 or dt == np.float16:
        return np.dtype(np.float32)
    return np.dtype(np.float64)


def _rows_last(a: np.ndarray, axis: int, dtype) -> tuple[np.ndarray, tuple]:
    """Move `axis` last, make C-contiguous (rows, n) of `dtype`."""
    moved = np.moveaxis(np.asarray(a), axis, -1)
    shape = moved.shape
    flat = np.ascontiguousarray(moved, dtype=dtype).reshape(-1, shape[-1])
    return flat, shape


def _restore(out2d: np.ndarray, shape: tuple, axis: int) -> np.ndarray:
    return np.moveaxis(out2d.reshape(shape), -1, axis)


# ----------------------------------------------------------------------------- moving
def _move(name: str, *arrs, window: int, min_count=None, axis: int = -1):
    if min_count is None:
        min_count = window
    dt = _float_loop_dtype(*arrs)
    flats = []
    shape = None
    for a in arrs:
        f, shape = _rows_last(a, axis, dt)
        flats.append(f)
    out = np.empty_like(flats[0])
    rows, n = flats[0].shape
    fn = getattr(lib(), f"orc_{name}_{_SUFFIX[dt]}")
    fn.restype = None
    if n > 0 and rows > 0:
        fn(*[_ptr(f) for f in flats], _ptr(out), _i64(rows), _i64(n), _i64(window), _i64(min_count))
    return _restore(out, shape, axis)


def move_mean(a, *, window, min_count=None, axis=-1):
    return _move("move_mean", a, window=window, min_count=min_count, axis=axis)


def move_sum(a, *, window, min_count=None, axis=-1):
    return _move("move_sum", a, window=window, min_count=min_count, axis=axis)


def move_std(a, *, window, min_count=None, axis=-1):
    return _move("move_std", a, window=window, min_count=min_count, axis=axis)


def move_var(a, *, window, min_count=None, axis=-1):
    return _move("move_var", a, window=window, min_count=min_count, axis=axis)


def move_cov(a, b, *, window, min_count=None, axis=-1):
    return _move("move_cov", a, b, window=window, min_count=min_count, axis=axis)


def move_corr(a, b, *, window, min_count=None, axis=-1):
    return _move("move_corr", a, b, window=window, min_count=min_count, axis=axis)


# ------------------------------------------------------------------------- exp moving
def _move_exp(name: str, *arrs, alpha, min_weight=0, axis: int = -1):
    arrs = [np.asarray(a) for a in arrs]
    n = arrs[0].shape[axis]
    if not isinstance(alpha, np.ndarray):
        # np.broadcast_to(python float) is float64: a float32 array then runs the float64
        # loop; np.float32 scalars keep float32.
        alpha_arr = np.broadcast_to(np.asarray(alpha), (n,))
    else:
        alpha_arr = alpha
    dt = _float_loop_dtype(*arrs, alpha_arr)
    flats = []
    shape = None
    for a in arrs:
        f, shape = _rows_last(a, axis, dt)
        flats.append(f)
    rows, n = flats[0].shape
    if alpha_arr.ndim <= 1:
        al = np.ascontiguousarray(np.broadcast_to(alpha_arr, (n,)), dtype=dt)
        alpha_rs = 0
    else:
        al, _ = _rows_last(np.broadcast_to(alpha_arr, arrs[0].shape), axis, dt)
        alpha_rs = n
    out = np.empty_like(flats[0])
    fn = getattr(lib(), f"orc_{name}_{_SUFFIX[dt]}")
    fn.restype = None
    mw = ctypes.c_float(min_weight) if dt == np.float32 else ctypes.c_double(min_weight)
    if n > 0 and rows > 0:
        fn(*[_ptr(f) for f in flats], _ptr(al), _i64(alpha_rs), mw, _ptr(out), _i64(rows), _i64(n))
    return _restore(out, shape, axis)


def move_exp_nancount(a, *, alpha, min_weight=0, axis=-1):
    return _move_exp("move_exp_nancount", a, alpha=alpha, min_weight=min_weight, axis=axis)


def move_exp_nanmean(a, *, alpha, min_weight=0, axis=-1):
    return _move_exp("move_exp_nanmean", a, alpha=alpha, min_weight=min_weight, axis=axis)


def move_exp_nansum(a, *, alpha, min_weight=0, axis=-1):
    return _move_exp("move_exp_nansum", a, alpha=alpha, min_weight=min_weight, axis=axis)


def move_exp_nanvar(a, *, alpha, min_weight=0, axis=-1):
    return _move_exp("move_exp_nanvar", a, alpha=alpha, min_weight=min_weight, axis=axis)


def move_exp_nanstd(a, *, alpha, min_weight=0, axis=-1):
    return _move_exp("move_exp_nanstd", a, alpha=alpha, min_weight=min_weight, axis=axis)


def move_exp_nancov(a, b, *, alpha, min_weight=0, axis=-1):
    return _move_exp("move_exp_nancov", a, b, alpha=alpha, min_weight=min_weight, axis=axis)


def move_exp_nancorr(a, b, *, alpha, min_weight=0, axis=-1):
    return _move_exp("move_exp_nancorr", a, b, alpha=alpha, min_weight=min_weight, axis=axis)


# ------------------------------------------------------------------------------ fills
def _fill(name: str, a, *, limit=None, axis=-1):
    a = np.asarray(a)
    if limit is None:
        limit = a.shape[axis]
    if a.dtype.kind in "iu":
        return a.copy()  # np.isnan is constant-false for integers: identity
    dt = a.dtype if a.dtype in _SUFFIX else _float_loop_dtype(a)
    flat, shape = _rows_last(a, axis, dt)
    out = np.empty_like(flat)
    rows, n = flat.shape
    fn = getattr(lib(), f"orc_{name}_{_SUFFIX[np.dtype(dt)]}")
    fn.restype = None
    if n > 0 and rows > 0:
        fn(_ptr(flat), _ptr(out), _i64(rows), _i64(n), _i64(limit))
    return _restore(out, shape, axis).astype(a.dtype, copy=False)


def ffill(a, *, limit=None, axis=-1):
    return _fill("ffill", a, limit=limit, axis=axis)


def bfill(a, *, limit=None, axis=-1):
    return _fill("bfill", a, limit=limit, axis=axis)


# ---------------------------------------------------------------------------- grouped
_FLOAT_ONLY = {"group_nanmean", "group_nanvar", "group_nanstd"}


def _group(name: str, values, labels, *, ddof=1, num_labels=None, axis=None):
    values = np.asarray(values)
    labels = np.asarray(labels)
    if values.dtype == np.bool_:
        values = values.astype(np.int32)
    if num_labels is None:
        num_labels = int(labels.max()) + 1
    if name in _FLOAT_ONLY and values.dtype.kind in "iu":
        vdt = np.dtype(np.float64)
    elif values.dtype in _SUFFIX:
        vdt = values.dtype
    elif values.dtype.kind in "iu":
        vdt = np.dtype(np.int64)
    else:
        vdt = _float_loop_dtype(values)
    ldt = np.dtype(np.int64) if labels.dtype.itemsize > 4 or values.size > np.iinfo(np.int32).max else np.dtype(np.int32)

    if axis is None:
        core = values.ndim
    elif isinstance(axis, int):
        values = np.moveaxis(values, axis, -1)
        core = 1
    else:
        values = np.moveaxis(values, axis, range(-len(axis), 0, 1))
        core = len(axis)
    bshape = values.shape[: values.ndim - core]
    n = int(np.prod(values.shape[values.ndim - core :], dtype=np.int64))
    v2 = np.ascontiguousarray(values, dtype=vdt).reshape(-1, n)
    l1 = np.ascontiguousarray(labels, dtype=ldt).reshape(-1)
    rows = v2.shape[0]
    out = np.empty((rows, num_labels), dtype=vdt)
    fn = getattr(lib(), f"orc_{name}_{_SUFFIX[vdt]}_{_SUFFIX[ldt]}")
    fn.restype = None
    if rows > 0:
        fn(_ptr(v2), _ptr(l1), _i64(0), _ptr(out), _i64(rows), _i64(n), _i64(num_labels), _i64(ddof))
    res = out.reshape(bshape + (num_labels,))
    if vdt != values.dtype and name not in _FLOAT_ONLY:
        res = res.astype(values.dtype)  # narrow ints: wrap-around like in-dtype accumulation
    return res


def _mk_group(name):
    def f(values, labels, *, ddof=1, num_labels=None, axis=None):
        return _group(name, values, labels, ddof=ddof, num_labels=num_labels, axis=axis)

    f.__name__ = name
    return f


group_nanmean = _mk_group("group_nanmean")
group_nansum = _mk_group("group_nansum")
group_nancount = _mk_group("group_nancount")
group_nanargmax = _mk_group("group_nanargmax")
group_nanargmin = _mk_group("group_nanargmin")
group_nanfirst = _mk_group("group_nanfirst")
group_nanlast = _mk_group("group_nanlast")
group_nanprod = _mk_group("group_nanprod")
group_nansum_of_squares = _mk_group("group_nansum_of_squares")
group_nanvar = _mk_group("group_nanvar")
group_nanstd = _mk_group("group_nanstd")
group_nanmin = _mk_group("group_nanmin")
group_nanmax = _mk_group("group_nanmax")
group_nanany = _mk_group("group_nanany")
group_nanall = _mk_group("group_nanall")

MOVE_FUNCS = ["move_mean", "move_sum", "move_std", "move_var", "move_cov", "move_corr"]
MOVE_EXP_FUNCS = [
    "move_exp_nancount",
    "move_exp_nanmean",
    "move_exp_nansum",
    "move_exp_nanvar",
    "move_exp_nanstd",
    "move_exp_nancov",
    "move_exp_nancorr",
]
FILL_FUNCS = ["ffill", "bfill"]
GROUPED_FUNCS = [
    "group_nanmean",
    "group_nansum",
    "group_nancount",
    "group_nanargmax",
    "group_nanargmin",
    "group_nanfirst",
    "group_nanlast",
    "group_nanprod",
    "group_nansum_of_squares",
    "group_nanvar",
    "group_nanstd",
    "group_nanmin",
    "group_nanmax",
    "group_nanany",
    "group_nanall",
]


# -------------------------------------------------- plain reductions
def _reduce_rows(a, axis, by_stride):
    """Move the reduced axes last (numbagg.utils.move_axes / np.moveaxis) and flatten them.
    by_stride: ndaggregate first sorts the axes by stride, largest first
    (decorators.py:211-230 _optimize_axis_order); ndreduce keeps the caller's order."""
    a = np.asarray(a)
    if axis is None:
        axis = tuple(range(a.ndim))
    elif not isinstance(axis, tuple):
        axis = (axis,)
    axis = tuple(ax % a.ndim for ax in axis)
    if by_stride and len(axis) > 1:
        order = np.argsort([a.strides[ax] for ax in axis])[::-1]
        axis = tuple(axis[i] for i in order)
    moved = np.moveaxis(a, axis, range(a.ndim - len(axis), a.ndim))
    bshape = moved.shape[: a.ndim - len(axis)]
    n = int(np.prod(moved.shape[a.ndim - len(axis):], dtype=np.int64))
    return moved, bshape, n


_EMPTY_ERRORS = {
    "nanargmax": "All-NaN slice encountered",
    "nanargmin": "All-NaN slice encountered",
    "nanmax": "zero-size array to reduction operation fmax which has no identity",
    "nanmin": "zero-size array to reduction operation fmin which has no identity",
}


def _reduce(name, a, out_dtype, *, axis=None, extra=(), by_stride=True):
    moved, bshape, n = _reduce_rows(a, axis, by_stride)
    dt = moved.dtype
    if dt == np.float16:
        dt = np.dtype(np.float32)
    if dt not in _SUFFIX:
        dt = np.dtype(np.float64) if dt.kind == "f" else np.dtype(np.int64)
    rows = int(np.prod(bshape, dtype=np.int64))
    if rows > 0 and n == 0 and name in _EMPTY_ERRORS:  # funcs.py:168-169,205-208: `if not a.size: raise`
        raise ValueError(_EMPTY_ERRORS[name])
    flat = np.ascontiguousarray(moved, dtype=dt).reshape(rows, n)
    out = np.empty(rows, dtype=out_dtype(dt) if isinstance(out_dtype, type(lambda: 0)) else out_dtype)
    fn = getattr(lib(), f"orc_{name}_{_SUFFIX[dt]}")
    fn.restype = None
    if rows > 0:
        fn(_ptr(flat), _ptr(out), _i64(rows), _i64(n), *extra)
    return out.reshape(bshape)[()] if bshape == () else out.reshape(bshape)


def allnan(a, *, axis=None):
    return _reduce("allnan", a, np.uint8, axis=axis).astype(np.bool_)


def anynan(a, *, axis=None):
    return _reduce("anynan", a, np.uint8, axis=axis).astype(np.bool_)


def nancount(a, *, axis=None):
    return _reduce("nancount", a, np.int64, axis=axis)


def nansum(a, *, axis=None):
    return _reduce("nansum", a, lambda dt: dt, axis=axis)


def nanmean(a, *, axis=None):
    return _reduce("nanmean", np.asarray(a, dtype=_float_loop_dtype(a)), lambda dt: dt, axis=axis)


def nanvar(a, *, ddof=1, axis=None):
    return _reduce("nanvarstd", np.asarray(a, dtype=_float_loop_dtype(a)), lambda dt: dt, axis=axis,
                   extra=(_i64(ddof), ctypes.c_int(0)))


def nanstd(a, *, ddof=1, axis=None):
    return _reduce("nanvarstd", np.asarray(a, dtype=_float_loop_dtype(a)), lambda dt: dt, axis=axis,
                   extra=(_i64(ddof), ctypes.c_int(1)))


def _arg(name, a, axis):
    res = _reduce(name, a, np.int64, axis=axis, by_stride=False)
    if np.any(np.asarray(res) < 0):
        raise ValueError("All-NaN slice encountered")
    return res


def nanargmax(a, *, axis=None):
    return _arg("nanargmax", a, axis)


def nanargmin(a, *, axis=None):
    return _arg("nanargmin", a, axis)


def nanmax(a, *, axis=None):
    return _reduce("nanmax", a, lambda dt: dt if dt.kind == "f" else np.dtype(np.int64), axis=axis,
                   by_stride=False)


def nanmin(a, *, axis=None):
    return _reduce("nanmin", a, lambda dt: dt if dt.kind == "f" else np.dtype(np.int64), axis=axis,
                   by_stride=False)


AGGREGATION_FUNCS = ["allnan", "anynan", "nancount", "nansum", "nanmean", "nanvar", "nanstd",
                     "nanargmax", "nanargmin", "nanmax", "nanmin"]


# ----------------------------------------------------------- quantiles
def nanquantile(a, quantiles, axis=None):
    """numbagg.nanquantile (funcs.py:245-291 behind ndquantile, decorators.py:821-884):
    float64 only; NaN = missing; linear interpolation between the two order statistics
    around rank (valid - 1) * q, with the reference's own arithmetic (rank, proportion and
    floor + proportion * (ceil - floor), each rounded separately).  The order statistics come
    from a full sort here (np.partition in the reference: same values)."""
    from collections.abc import Iterable

    squeeze = not isinstance(quantiles, Iterable)
    q = np.asarray([quantiles] if squeeze else quantiles, dtype=np.float64)
    if np.any(q < 0) or np.any(q > 1):
        raise ValueError(f"quantiles must be in the range [0, 1], inclusive. Got {q}.")
    a = np.asarray(a)
    if axis is None:
        axis = tuple(range(a.ndim))
    elif not isinstance(axis, tuple):
        axis = (axis,)
    moved, bshape, n = _reduce_rows(a, axis, by_stride=False)
    rows = int(np.prod(bshape, dtype=np.int64))
    flat = np.ascontiguousarray(moved, dtype=np.float64).reshape(rows, n)
    out = np.empty((rows, len(q)), dtype=np.float64)
    for r in range(rows):
        x = flat[r]
        valid = int(n - np.count_nonzero(np.isnan(x)))
        if valid == 0:
            out[r] = np.nan
            continue
        s = np.sort(x)  # NaN sorts last, like the reference's fill with the maximum
        for i, qi in enumerate(q):
            if np.isnan(qi):
                out[r, i] = np.nan
                continue
            rank = np.float64(valid - 1) * qi
            lo, hi = int(np.floor(rank)), int(np.ceil(rank))
            proportion = rank - np.float64(lo)
            with np.errstate(invalid="ignore"):
                out[r, i] = s[lo] + proportion * (s[hi] - s[lo])
    res = np.moveaxis(out.reshape(bshape + (len(q),)), -1, 0)
    return res[0] if squeeze else res


def nanmedian(a, axis=None):
    return nanquantile(a, 0.5, axis=axis)


# -------------------------------------------- matrix functions (oracle only)
def _matrix_input(a, what):
    a = np.asarray(a)
    dt = _float_loop_dtype(a)
    return np.ascontiguousarray(a, dtype=dt), dt


def _static_matrix(a, corr, name):
    if np.asarray(a).ndim < 2:
        raise ValueError(
            f"{name} requires at least a 2D array with shape (..., vars, obs). "
            "For 1D arrays, use nanvar for variance calculations."
        )
    a, dt = _matrix_input(a, name)
    nv, no = a.shape[-2:]
    batch = int(np.prod(a.shape[:-2], dtype=np.int64))
    out = np.empty(a.shape[:-2] + (nv, nv), dtype=dt)
    fn = getattr(lib(), f"orc_nanmatrix_{_SUFFIX[dt]}")
    fn.restype = None
    if batch > 0 and nv > 0:
        fn(_ptr(a), _ptr(out), _i64(batch), _i64(nv), _i64(no), ctypes.c_int(corr))
    return out


def nancorrmatrix(a):
    return _static_matrix(a, 1, "nancorrmatrix")


def nancovmatrix(a):
    return _static_matrix(a, 0, "nancovmatrix")


def _move_matrix(a, window, min_count, corr, name):
    a = np.asarray(a)
    if a.ndim < 2:
        raise ValueError(f"{name} requires at least a 2D array with shape (..., obs, vars).")
    if min_count is None:
        min_count = window
    elif min_count < 0:
        raise ValueError(f"min_count must be positive: {min_count}")
    if not 0 < window <= a.shape[-2]:
        raise ValueError(f"window not in valid range: {window}")
    a, dt = _matrix_input(a, name)
    no, nv = a.shape[-2:]
    batch = int(np.prod(a.shape[:-2], dtype=np.int64))
    out = np.empty(a.shape[:-2] + (no, nv, nv), dtype=dt)
    fn = getattr(lib(), f"orc_move_matrix_{_SUFFIX[dt]}")
    fn.restype = None
    if batch > 0 and nv > 0:
        fn(_ptr(a), _ptr(out), _i64(batch), _i64(no), _i64(nv), _i64(window), _i64(min_count), ctypes.c_int(corr))
    return out


def move_corrmatrix(a, window, min_count=None):
    return _move_matrix(a, window, min_count, 1, "move_corrmatrix")


def move_covmatrix(a, window, min_count=None):
    return _move_matrix(a, window, min_count, 0, "move_covmatrix")


def _move_exp_matrix(a, alpha, min_weight, corr, name):
    a = np.asarray(a)
    if a.ndim < 2:
        raise ValueError(f"{name} requires at least a 2D array with shape (..., obs, vars).")
    if not isinstance(alpha, np.ndarray):
        alpha = np.broadcast_to(alpha, a.shape[-2])
    # gufunc loop selection over (a, alpha); a Python-scalar min_weight is weakly typed
    dt = _float_loop_dtype(a, alpha) if not isinstance(min_weight, np.generic) else _float_loop_dtype(a, alpha, np.asarray(min_weight))
    a = np.ascontiguousarray(a, dtype=dt)
    no, nv = a.shape[-2:]
    bshape = a.shape[:-2]
    batch = int(np.prod(bshape, dtype=np.int64))
    alpha = np.asarray(alpha, dtype=dt)
    if alpha.ndim == 1:
        per_item, al = 0, np.ascontiguousarray(alpha)
    else:
        per_item, al = 1, np.ascontiguousarray(np.broadcast_to(alpha, bshape + (no,)))
    out = np.empty(bshape + (no, nv, nv), dtype=dt)
    fn = getattr(lib(), f"orc_move_exp_matrix_{_SUFFIX[dt]}")
    fn.restype = None
    if batch > 0 and nv > 0:
        fn(_ptr(a), _ptr(al), ctypes.c_int(per_item), ctypes.c_double(float(min_weight)), _ptr(out),
           _i64(batch), _i64(no), _i64(nv), ctypes.c_int(corr))
    return out


def move_exp_nancorrmatrix(a, alpha, min_weight=0):
    return _move_exp_matrix(a, alpha, min_weight, 1, "move_exp_nancorrmatrix")


def move_exp_nancovmatrix(a, alpha, min_weight=0):
    return _move_exp_matrix(a, alpha, min_weight, 0, "move_exp_nancovmatrix")


MATRIX_FUNCS = ["nancorrmatrix", "nancovmatrix", "move_corrmatrix", "move_covmatrix",
                "move_exp_nancorrmatrix", "move_exp_nancovmatrix"]
