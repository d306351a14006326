"""Pin the C oracle (oracle/nbg_oracle.c) to the reference: every golden case produced by
numbagg's own Numba path (oracle/gen_golden.py) must be reproduced BIT-FOR-BIT -- the oracle
restates the same sequential loops with the same types, so no tolerance is needed.
CPU-only; runs in the dev container and on the GPU box."""

import numpy as np
import pytest

from oracle import oracle
from tests._golden import all_cases

CASES = all_cases() + all_cases("reduce") + all_cases("quantile") + all_cases("matrix")


def _int_empty_mask(case, expected):
    """Integer outputs for empty groups are uninitialised / NaN-cast in the reference
    (grouped.py:95-110): exclude those slots from the comparison."""
    if not case.func.startswith("group_") or expected.dtype.kind not in "iu":
        return None
    if case.func not in ("group_nanfirst", "group_nanlast", "group_nanargmax", "group_nanargmin",
                         "group_nanmin", "group_nanmax"):
        return None
    counts = oracle.group_nancount(*case.args, **case.kwargs)
    return counts == 0


@pytest.mark.parametrize("case", CASES, ids=[c.id for c in CASES])
def test_oracle_matches_reference_bits(case):
    got = getattr(oracle, case.func)(*case.args, **case.kwargs)
    exp = case.expected
    got = np.asarray(got)
    assert got.shape == exp.shape
    assert got.dtype == exp.dtype
    mask = _int_empty_mask(case, exp)
    if mask is not None:
        got = np.where(mask, 0, got)
        exp = np.where(mask, 0, exp)
    # assert_array_equal treats NaN == NaN and +0 == -0; additionally require identical
    # NaN masks and, for floats, identical bit patterns of every finite value.
    np.testing.assert_array_equal(got, exp)
    if exp.dtype.kind == "f":
        fin = np.isfinite(exp)
        assert np.array_equal(np.isnan(got), np.isnan(exp))
        assert np.array_equal(got[fin].view(f"u{exp.dtype.itemsize}"), exp[fin].view(f"u{exp.dtype.itemsize}"))
