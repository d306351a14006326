"""Parity helpers implementing the north_star tolerances (BASELINE.json):
bit-exact for ffill/bfill, counts, min/max, arg*, first/last, any/all; rtol 1e-12 (float64) or
1e-5 (float32) for sums, means, var/std, cov/corr and the exponential moving functions.  NaN
masks must match exactly in both classes."""

import numpy as np

from tests._golden import EXACT_FUNCS

RTOL = {np.dtype(np.float64): 1e-12, np.dtype(np.float32): 1e-5}

# Observed worst relative error per (function, dtype) over everything the session compared in the
# tolerance class -- written to gpurun_out/parity_observed.json by tests/conftest.py so that the
# floors below can be tightened with data and regressions inside them stay visible.
OBSERVED: dict = {}


def _record(func, got, exp, rtol, atol):
    g = np.asarray(got, dtype=np.float64)
    e = np.asarray(exp, dtype=np.float64)
    ok = np.isfinite(g) & np.isfinite(e)
    if not ok.any():
        return
    err = np.abs(g[ok] - e[ok])
    rel = float(np.max(err / np.maximum(np.abs(e[ok]), 1e-300)))
    # error in units of the allowed bound (1.0 = at the limit)
    used = float(np.max(err / (atol + rtol * np.abs(e[ok]) + 1e-300)))
    key = f"{func}:{np.asarray(exp).dtype}"
    o = OBSERVED.setdefault(key, dict(max_rel=0.0, max_frac_of_bound=0.0, rtol=rtol, cases=0))
    o["max_rel"] = max(o["max_rel"], rel)
    o["max_frac_of_bound"] = max(o["max_frac_of_bound"], used)
    o["cases"] += 1


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
        _record(func + "^2", got.astype(np.float64) ** 2, exp.astype(np.float64) ** 2, 2 * rt, rt * scale)
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
    _record(func, got, exp, rtol, atol)
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
