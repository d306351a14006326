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
    """Magnitude of the sums a cancellation-prone output is formed from."""
    f = case.func
    if f in ("move_var", "move_std", "move_cov", "move_corr", "move_exp_nanvar", "move_exp_nanstd",
             "move_exp_nancov", "move_exp_nancorr", "group_nanvar", "group_nanstd"):
        m = max(float(np.nanmax(np.abs(np.where(np.isfinite(a), a, 0.0)), initial=0.0))
                for a in case.args if np.asarray(a).dtype.kind == "f")
        return m * m if "corr" not in f else 1.0
    return None


@pytest.mark.parametrize("case", CASES, ids=[c.id for c in CASES])
def test_cuda_matches_reference(case):
    import numbagg_b200 as nb

    got = getattr(nb, case.func)(*case.args, **case.kwargs)
    mask = int_empty_mask(case.func, case.args, case.kwargs, case.expected, oracle)
    assert_parity(case.func, got, case.expected, scale=_scale(case), int_empty_mask=mask)
