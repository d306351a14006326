"""Matrix functions on the GPU (nbg_matrix) against the reference's frozen outputs
(tests/golden/matrix.npz) and the oracle.  The kernels run the reference's loop body per
(i, j) pair in the reference's order and types, so every comparison is BIT-EXACT for float32
and float64 alike (NaN masks included)."""

import numpy as np
import pytest

from oracle import oracle
from tests._golden import all_cases

pytestmark = pytest.mark.gpu

CASES = all_cases("matrix")


def same(got, exp):
    got, exp = np.asarray(got), np.asarray(exp)
    assert got.shape == exp.shape, (got.shape, exp.shape)
    assert got.dtype == exp.dtype, (got.dtype, exp.dtype)
    np.testing.assert_array_equal(got, exp)
    if exp.dtype.kind == "f":
        fin = np.isfinite(exp)
        assert np.array_equal(got[fin].view(f"u{exp.dtype.itemsize}"), exp[fin].view(f"u{exp.dtype.itemsize}"))


@pytest.mark.parametrize("case", CASES, ids=[c.id for c in CASES])
def test_cuda_matches_reference(case):
    import numbagg_b200 as nb

    same(getattr(nb, case.func)(*case.args, **case.kwargs), case.expected)


def _data(shape, dtype, seed, nan_frac=0.15):
    rs = np.random.RandomState(seed)
    a = rs.standard_normal(shape) + 3.0
    a[rs.rand(*shape) < nan_frac] = np.nan
    return a.astype(dtype)


@pytest.mark.parametrize("dtype", [np.float64, np.float32])
def test_larger_shapes_against_oracle(dtype):
    import numbagg_b200 as nb

    vo = _data((3, 40, 700), dtype, seed=1)       # (batch, vars, obs)
    same(nb.nancorrmatrix(vo), oracle.nancorrmatrix(vo))
    same(nb.nancovmatrix(vo), oracle.nancovmatrix(vo))
    ov = _data((2, 3000, 12), dtype, seed=2)      # (batch, obs, vars)
    for w, mc in ((50, None), (500, 20)):
        same(nb.move_corrmatrix(ov, window=w, min_count=mc), oracle.move_corrmatrix(ov, window=w, min_count=mc))
        same(nb.move_covmatrix(ov, window=w, min_count=mc), oracle.move_covmatrix(ov, window=w, min_count=mc))
    al = (np.random.RandomState(3).rand(3000) * 0.5 + 0.01).astype(dtype)
    for alpha in (dtype(0.05), al, np.broadcast_to(al, (2, 3000)).copy()):
        same(nb.move_exp_nancorrmatrix(ov, alpha=alpha, min_weight=0.1),
             oracle.move_exp_nancorrmatrix(ov, alpha=alpha, min_weight=0.1))
        same(nb.move_exp_nancovmatrix(ov, alpha=alpha), oracle.move_exp_nancovmatrix(ov, alpha=alpha))


@pytest.mark.parametrize("dtype", [np.float64, np.float32])
def test_long_observation_axis_segments(dtype):
    """Observation axes >= 8192 are cut into segments that rebuild their window state from the
    `window` observations before them (nbg_matrix.cu: mat_move_seg_kernel).  The reference never
    re-syncs its running sums (moving_matrix.py:78-101), so the comparison is to the rounding of those
    sums, kept in the INPUT dtype: ~sqrt(steps) ulps of a window sum (values ~3, window 64: float64
    sums ~600 -> 1e-11 absolute on a covariance; float32 -> 1e-3), NaN masks exact."""
    import numbagg_b200 as nb
    from tests._parity import _record

    ov = _data((2, 20_000, 8), dtype, seed=5)
    atol = 1e-10 if dtype is np.float64 else 3e-3
    for func, kw in (("move_covmatrix", dict(window=64, min_count=8)), ("move_corrmatrix", dict(window=64, min_count=8)),
                     ("move_covmatrix", dict(window=5000, min_count=None))):
        got = getattr(nb, func)(ov, **kw)
        exp = getattr(oracle, func)(ov, **kw)
        assert got.dtype == exp.dtype and got.shape == exp.shape
        assert np.array_equal(np.isnan(got), np.isnan(exp)), f"{func}: NaN masks differ"
        _record(func + "(segments)", got, exp, 0.0, atol)
        np.testing.assert_allclose(got, exp, rtol=0, atol=atol, equal_nan=True)
    # other pair counts: 40 and 130 variables
    for nvars in (40, 130):
        ow = _data((1, 8500, nvars), dtype, seed=7)
        got = nb.move_covmatrix(ow, window=30, min_count=3)
        exp = oracle.move_covmatrix(ow, window=30, min_count=3)
        assert np.array_equal(np.isnan(got), np.isnan(exp))
        np.testing.assert_allclose(got, exp, rtol=0, atol=atol, equal_nan=True)
    # exponential weights: segments carry their state as an affine map (two passes)
    al = (np.random.RandomState(6).rand(20_000) * 0.2 + 0.005).astype(dtype)
    for alpha in (dtype(0.02), al):
        for func, kw in (("move_exp_nancovmatrix", dict(alpha=alpha)), ("move_exp_nancorrmatrix", dict(alpha=alpha, min_weight=0.2))):
            got = getattr(nb, func)(ov, **kw)
            exp = getattr(oracle, func)(ov, **kw)
            assert got.dtype == exp.dtype and got.shape == exp.shape
            assert np.array_equal(np.isnan(got), np.isnan(exp)), f"{func}: NaN masks differ"
            _record(func + "(segments)", got, exp, 0.0, atol)
            np.testing.assert_allclose(got, exp, rtol=0, atol=atol, equal_nan=True)
    # static forms: partial sums per segment of the observation axis, folded in order
    vo = np.ascontiguousarray(np.swapaxes(ov, -1, -2))  # (batch, vars, obs)
    for func in ("nancovmatrix", "nancorrmatrix"):
        got = getattr(nb, func)(vo)
        exp = getattr(oracle, func)(vo)
        assert got.dtype == exp.dtype and got.shape == exp.shape
        assert np.array_equal(np.isnan(got), np.isnan(exp)), f"{func}: NaN masks differ"
        _record(func + "(segments)", got, exp, 0.0, atol)
        np.testing.assert_allclose(got, exp, rtol=0, atol=atol, equal_nan=True)
    # the first segment runs the reference's recurrence from the start: bit-identical there
    got = nb.move_covmatrix(ov, window=64, min_count=8)
    same(got[:, :256], oracle.move_covmatrix(ov, window=64, min_count=8)[:, :256])


def test_tensor_in_tensor_out_and_validation():
    import torch

    import numbagg_b200 as nb

    ov = _data((200, 6), np.float64, seed=4)
    t = torch.from_numpy(ov).cuda()
    got = nb.move_covmatrix(t, window=20, min_count=5)
    assert isinstance(got, torch.Tensor) and got.is_cuda
    same(got.cpu().numpy(), oracle.move_covmatrix(ov, window=20, min_count=5))
    same(nb.nancorrmatrix(t.T.contiguous()).cpu().numpy(), oracle.nancorrmatrix(np.ascontiguousarray(ov.T)))
    with pytest.raises(ValueError, match="requires at least a 2D array"):
        nb.nancorrmatrix(np.arange(5.0))
    with pytest.raises(ValueError, match="requires at least a 2D array"):
        nb.move_corrmatrix(np.arange(5.0), window=2)
    with pytest.raises(ValueError, match="window not in valid range"):
        nb.move_corrmatrix(ov, window=0)
    with pytest.raises(ValueError, match="window not in valid range"):
        nb.move_covmatrix(ov, window=201)
    with pytest.raises(ValueError, match="min_count must be positive"):
        nb.move_covmatrix(ov, window=5, min_count=-1)
    # a Python-float alpha promotes float32 data to the float64 loop, np.float32 does not
    o32 = ov.astype(np.float32)
    assert nb.move_exp_nancovmatrix(o32, alpha=0.1).dtype == np.float64
    assert nb.move_exp_nancovmatrix(o32, alpha=np.float32(0.1)).dtype == np.float32
