"""Parity helpers implementing the north_star tolerances (BASELINE.json):
bit-exact for ffill/bfill, counts, min/max, arg*, first/last, any/all; rtol 1e-12 (float64) or
1e-5 (float32) for sums, means, var/std, cov/corr and the exponential moving functions.  NaN
masks must match exactly in both classes."""

import numpy as np

from tests._golden import EXACT_FUNCS

RTOL = {np.dtype(np.float64): 1e-12, np.dtype(np.float32): 1e-5}


def assert_parity(func: str, got, exp, *, scale=None, int_empty_mask=None, atol=None):
    """`scale`: magnitude of the sums the output is a difference of (absolute floor =
    rtol*scale).  Standard deviations are compared as variances: the floor belongs to the
    variance, and sqrt amplifies it by 1/(2*std) for near-constant windows.  `atol` overrides
    the floor for outputs that DIVIDE by such differences (correlations of tiny windows)."""
    got = np.asarray(got)
    exp = np.asarray(exp)
    if func.endswith("std") and np.asarray(exp).dtype.kind == "f" and scale is not None:
        assert np.array_equal(np.isnan(got), np.isnan(exp)), f"{func}: NaN masks differ"
        rt = RTOL[np.asarray(exp).dtype]
        np.testing.assert_allclose(got.astype(np.float64) ** 2, exp.astype(np.float64) ** 2,
                                   rtol=2 * rt, atol=rt * scale, equal_nan=True)
        return
    assert got.shape == exp.shape, (got.shape, exp.shape)
    assert got.dtype == exp.dtype, (got.dtype, exp.dtype)
    if int_empty_mask is not None:
        got = np.where(int_empty_mask, 0, got)
        exp = np.where(int_empty_mask, 0, exp)
    if exp.dtype.kind != "f" or func in EXACT_FUNCS:
        np.testing.assert_array_equal(got, exp)
        return
    assert np.array_equal(np.isnan(got), np.isnan(exp)), (
        f"{func}: NaN masks differ at {np.flatnonzero(np.isnan(got) != np.isnan(exp))[:10]}"
    )
    rtol = RTOL[exp.dtype]
    # absolute floor: outputs that are differences of O(scale) sums (cancellation) cannot be
    # relatively accurate to rtol in ANY summation order, the reference's included
    if atol is None:
        atol = rtol * (scale if scale is not None else 0.0)
    np.testing.assert_allclose(got, exp, rtol=rtol, atol=atol, equal_nan=True)


def int_empty_mask(func, args, kwargs, exp, oracle):
    """Integer outputs for empty groups are uninitialised / NaN-cast in the reference
    (grouped.py:95-110): exclude those slots."""
    if not func.startswith("group_") or np.asarray(exp).dtype.kind not in "iu":
        return None
    if func not in ("group_nanfirst", "group_nanlast", "group_nanargmax", "group_nanargmin",
                    "group_nanmin", "group_nanmax"):
        return None
    return oracle.group_nancount(*args, **kwargs) == 0
