"""CUDA path vs the reference's own outputs (tests/golden, produced by numbagg's Numba path)
through the public drop-in API with numpy inputs.  Tolerances: tests/_parity.py."""

import numpy as np
import pytest

from oracle import oracle
from tests._golden import all_cases
from tests._parity import assert_parity, int_empty_mask

pytestmark = pytest.mark.gpu

CASES = all_cases()


def _scale(case):
    """Magnitude of the running sums an output is formed from.  Outputs that are differences
    of O(scale) sums (an emptied window, a variance) carry an absolute rounding residue of a
    few ulps of `scale` in ANY summation order -- the reference's own drift included (it never
    re-syncs its running sums, moving.py:30-55) -- so the relative tolerance gets an absolute
    floor of rtol*scale."""
    f = case.func
    floats = [np.asarray(a, dtype=np.float64) for a in case.args if np.asarray(a).dtype.kind in "fiub"]
    if not floats:
        return None
    m = max(float(np.max(np.abs(np.where(np.isfinite(a), a, 0.0)), initial=0.0)) for a in floats)
    second_moment = any(t in f for t in ("var", "std", "cov"))
    if "corr" in f:
        return 1.0
    scale = m * m if second_moment else m
    if f == "move_sum":
        scale *= case.kwargs["window"]
    if f == "move_exp_nansum":
        alpha = case.kwargs["alpha"]
        scale /= max(float(np.min(alpha)), 1e-3)
    if f in ("group_nansum", "group_nansum_of_squares"):
        scale = None  # plain sums of same-sign data: relative tolerance only
    return scale


@pytest.mark.parametrize("case", CASES, ids=[c.id for c in CASES])
def test_cuda_matches_reference(case):
    import numbagg_b200 as nb

    got = getattr(nb, case.func)(*case.args, **case.kwargs)
    mask = int_empty_mask(case.func, case.args, case.kwargs, case.expected, oracle)
    assert_parity(case.func, got, case.expected, scale=_scale(case), int_empty_mask=mask)
