"""ctypes binding of libnbg_b200.so (include/nbg_b200.h).  No fallback: if the CUDA library
is missing or a call fails, the caller gets an exception."""

from __future__ import annotations

import ctypes
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libnbg_b200.so")

# enums of include/nbg_b200.h
NBG_F32, NBG_F64, NBG_I32, NBG_I64 = 0, 1, 2, 3
MOVE_OPS = dict(move_mean=0, move_sum=1, move_std=2, move_var=3, move_cov=4, move_corr=5)
EXP_OPS = dict(
    move_exp_nancount=0, move_exp_nanmean=1, move_exp_nansum=2, move_exp_nanvar=3,
    move_exp_nanstd=4, move_exp_nancov=5, move_exp_nancorr=6,
)
FILL_DIRS = dict(ffill=0, bfill=1)
GROUP_OPS = dict(
    group_nanmean=0, group_nansum=1, group_nancount=2, group_nanargmax=3, group_nanargmin=4,
    group_nanfirst=5, group_nanlast=6, group_nanprod=7, group_nansum_of_squares=8,
    group_nanvar=9, group_nanstd=10, group_nanmin=11, group_nanmax=12, group_nanany=13,
    group_nanall=14,
)
REDUCE_OPS = dict(
    allnan=0, anynan=1, nancount=2, nansum=3, nanmean=4, nanvar=5, nanstd=6, nanargmax=7,
    nanargmin=8, nanmax=9, nanmin=10,
)
MATRIX_OPS = dict(
    nancorrmatrix=0, nancovmatrix=1, move_corrmatrix=2, move_covmatrix=3, move_exp_nancorrmatrix=4,
    move_exp_nancovmatrix=5,
)
NBG_REDUCE_STATE_WORDS = 3
NBG_QUANTILE_MAX_Q = 16
NBG_EXP_STATE = 11
NBG_FILL_STATE = 3
NBG_GROUP_WS_CHANNELS = 4

_i64 = ctypes.c_int64
_vp = ctypes.c_void_p
_int = ctypes.c_int
_dbl = ctypes.c_double
_sz = ctypes.c_size_t

_SIGNATURES = {
    "nbg_abi_version": (_int, []),
    "nbg_last_error": (ctypes.c_char_p, []),
    "nbg_launch_count": (_i64, []),
    "nbg_move": (_int, [_int, _int, _vp, _vp, _vp, _i64, _i64, _i64, _i64, _i64, _vp, _vp, _i64, _vp]),
    "nbg_move_exp": (
        _int,
        [_int, _int, _vp, _vp, _vp, _int, _dbl, _dbl, _vp, _i64, _i64, _i64, _vp, _vp, _vp, _sz, _vp],
    ),
    "nbg_move_exp_workspace_bytes": (_sz, [_int, _int, _i64, _i64, _i64]),
    "nbg_fill": (_int, [_int, _int, _vp, _vp, _i64, _i64, _i64, _i64, _vp, _vp, _vp, _sz, _vp]),
    "nbg_fill_workspace_bytes": (_sz, [_int, _i64, _i64, _i64]),
    "nbg_fill_sentinel_bits": (ctypes.c_uint64, [_int]),
    "nbg_fill_patch": (_int, [_int, _int, _vp, _i64, _i64, _i64, _i64, _vp, _vp]),
    "nbg_group_record_words": (_int, [_int]),
    "nbg_group_record_layout": (_int, [_int, _i64, _i64, ctypes.POINTER(_i64)]),
    "nbg_group_workspace_bytes": (_sz, [_int, _int, _i64, _i64, _i64]),
    "nbg_group_init": (_int, [_int, _int, _vp, _i64, _i64, _vp]),
    "nbg_group_accumulate": (_int, [_int, _int, _int, _vp, _vp, _int, _vp, _sz, _i64, _i64, _i64, _i64, _vp]),
    "nbg_group_combine": (_int, [_int, _int, _vp, _vp, _i64, _i64, _vp]),
    "nbg_group_finalize": (_int, [_int, _int, _vp, _vp, _i64, _i64, _i64, _vp]),
    "nbg_group": (_int, [_int, _int, _int, _vp, _vp, _int, _vp, _i64, _i64, _i64, _i64, _vp, _sz, _vp]),
    "nbg_reduce_workspace_bytes": (_sz, [_int, _int, _i64, _i64, _i64]),
    "nbg_reduce": (_int, [_int, _int, _vp, _vp, _i64, _i64, _i64, _i64, _vp, _sz, _vp]),
    "nbg_reduce_partial": (_int, [_int, _int, _vp, _vp, _i64, _i64, _i64, _i64, _vp, _sz, _vp]),
    "nbg_reduce_merge": (_int, [_int, _int, _vp, _i64, _i64, _vp, _i64, _i64, _vp]),
    "nbg_quantile_workspace_bytes": (_sz, [_i64, _i64, _i64]),
    "nbg_quantile": (_int, [_vp, _vp, _vp, _i64, _i64, _i64, _vp, _sz, _vp]),
    "nbg_matrix": (_int, [_int, _int, _vp, _vp, _int, _dbl, _vp, _i64, _i64, _i64, _i64, _i64, _vp]),
}

EXPORTED_SYMBOLS = tuple(_SIGNATURES)

_lib = None


class NbgError(RuntimeError):
    pass


def lib() -> ctypes.CDLL:
    """Load the C-ABI library.  Raises (never falls back) when it has not been built."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise ImportError(
                f"{LIB_PATH} is missing: numbagg_b200 has no CPU fallback. Build the CUDA "
                "library first: `python -m numbagg_b200.build` (needs nvcc)."
            )
        handle = ctypes.CDLL(LIB_PATH)
        for name, (res, args) in _SIGNATURES.items():
            fn = getattr(handle, name)  # AttributeError if the .so lacks a declared symbol
            fn.restype = res
            fn.argtypes = args
        if handle.nbg_abi_version() != 1:
            raise ImportError("libnbg_b200.so ABI version mismatch; rebuild with `python -m numbagg_b200.build --force`")
        _lib = handle
    return _lib


def check(rc: int, what: str) -> None:
    if rc != 0:
        msg = lib().nbg_last_error().decode(errors="replace")
        raise NbgError(f"{what} failed (status {rc}): {msg}")


def launch_count() -> int:
    return int(lib().nbg_launch_count())
