"""nanquantile / nanmedian on the GPU (nbg_quantile: shared-memory sort for short rows, radix
select for long ones) against the reference's frozen outputs (tests/golden/quantile.npz) and
the oracle.  Selection is exact and the interpolation repeats the reference's arithmetic, so
every comparison is BIT-EXACT (NaN masks included; the sign of a zero is not compared)."""

import numpy as np
import pytest

from oracle import oracle
from tests._golden import all_cases

pytestmark = pytest.mark.gpu

CASES = all_cases("quantile")


def same(got, exp):
    got, exp = np.asarray(got), np.asarray(exp)
    assert got.shape == exp.shape, (got.shape, exp.shape)
    assert got.dtype == exp.dtype == np.float64
    np.testing.assert_array_equal(got, exp)


@pytest.mark.parametrize("case", CASES, ids=[c.id for c in CASES])
def test_cuda_matches_reference(case):
    import numbagg_b200 as nb

    same(getattr(nb, case.func)(*case.args, **case.kwargs), case.expected)


def _data(shape, seed, nan_frac=0.2, ties=False):
    rs = np.random.RandomState(seed)
    a = rs.standard_normal(shape) * 10.0 ** rs.randint(-3, 4)
    if ties:
        a = np.round(a)
    a[rs.rand(*shape) < nan_frac] = np.nan
    return a


SHAPES = [((300, 1), -1), ((200, 2), -1), ((100, 33), -1), ((50, 1000), -1), ((9, 4096), -1), ((7, 4097), -1),
          ((3, 100_001), -1), ((1, 1_000_000), -1), ((1_000_000,), None), ((40, 5000), -1), ((5000, 40), 0),
          ((12, 70, 50), (0, 2)), ((12, 70, 50), 1), ((6, 9000, 3), 1)]


@pytest.mark.parametrize("ties", [False, True])
@pytest.mark.parametrize("shape,axis", SHAPES, ids=[f"{s}-{a}" for s, a in SHAPES])
def test_sort_and_select_paths_against_oracle(shape, axis, ties):
    import numbagg_b200 as nb

    a = _data(shape, seed=len(shape) + shape[0] % 17, ties=ties)
    for q in (0.5, [0.0, 0.25, 0.5, 0.75, 1.0], [0.37, np.nan]):
        same(nb.nanquantile(a, q, axis=axis), oracle.nanquantile(a, q, axis=axis))
    same(nb.nanmedian(a, axis=axis), oracle.nanmedian(a, axis=axis))


def test_many_quantiles_dtypes_and_specials():
    import torch

    import numbagg_b200 as nb

    a = _data((20, 6000), seed=3)
    q = np.linspace(0, 1, 41)  # more than one 16-quantile call
    same(nb.nanquantile(a, q, axis=-1), oracle.nanquantile(a, q, axis=-1))
    a[3] = np.nan  # a row without data
    a[4, :3000] = np.inf
    a[5, ::2] = -np.inf
    a[6] = 7.25
    same(nb.nanquantile(a, [0.0, 0.5, 0.9, 1.0], axis=-1), oracle.nanquantile(a, [0.0, 0.5, 0.9, 1.0], axis=-1))
    for dt in (np.float32, np.int32, np.int64):
        x = (np.random.RandomState(5).standard_normal((30, 500)) * 100).astype(dt)
        same(nb.nanquantile(x, [0.1, 0.5], axis=0), oracle.nanquantile(x, [0.1, 0.5], axis=0))
    t = torch.from_numpy(a).cuda()
    got = nb.nanquantile(t, 0.3, axis=-1)
    assert isinstance(got, torch.Tensor) and got.is_cuda
    same(got.cpu().numpy(), oracle.nanquantile(a, 0.3, axis=-1))
    same(nb.nanquantile(np.empty((3, 0)), 0.5, axis=-1), oracle.nanquantile(np.empty((3, 0)), 0.5, axis=-1))


def test_validation_matches_reference():
    import numbagg_b200 as nb

    a = np.arange(10.0)
    for bad in (1.5, [-0.1, 0.5]):
        with pytest.raises(ValueError, match="quantiles must be in the range"):
            nb.nanquantile(a, bad)
    with pytest.raises(np.exceptions.AxisError):
        nb.nanquantile(a, 0.5, axis=1)
    assert nb.nanquantile(a, 0.45) == oracle.nanquantile(a, 0.45)


def test_full_size_properties():
    """10^8 float64 elements in one row (radix select over 800 MB): quantiles are monotone,
    q=0/1 are min/max, the median splits the data in two halves."""
    import torch

    import numbagg_b200 as nb

    g = torch.Generator(device="cuda").manual_seed(3)
    t = torch.randn(100_000_000, generator=g, device="cuda", dtype=torch.float64)
    t[::13] = float("nan")
    q = [0.0, 0.001, 0.25, 0.5, 0.75, 0.999, 1.0]
    r = nb.nanquantile(t, q)
    assert bool((r[1:] >= r[:-1]).all())
    assert float(r[0]) == float(nb.nanmin(t)) and float(r[-1]) == float(nb.nanmax(t))
    valid = int(nb.nancount(t))
    below = int((t < r[3]).sum())
    assert abs(below - valid / 2) <= 1
    assert abs(float(r[3])) < 1e-3 and abs(float(r[2]) + 0.6745) < 1e-3


def test_many_long_rows_are_batched():
    """More long rows than one launch of the radix-select path takes (the host splits them)."""
    import numbagg_b200 as nb

    a = _data((70_000, 4100), seed=11, nan_frac=0.1)
    got = nb.nanquantile(a, [0.5, 0.9], axis=-1)
    sub = np.r_[0:50, 32760:32780, 69_990:70_000]
    same(got[:, sub], oracle.nanquantile(a[sub], [0.5, 0.9], axis=-1))
