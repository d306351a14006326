"""SURVEY 8(f) rank 4 -- the integration surface the reference gets from NumPy's gufunc machinery:
`out=` / `dtype=` / `casting=` forwarded through **kwargs (numbagg/decorators.py:311-341, 380-414,
471-487, 712-731, 783-807, 836-876, 1069-1089), big-endian and narrow integer / bool / float16 inputs
(test/util.py:49-54), empty arrays.  Expected behaviour was probed against the reference itself
(numba 0.65, numpy 2.3) and is restated next to each assertion."""

import numpy as np
import pytest

from oracle import oracle

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def nb():
    import numbagg_b200

    return numbagg_b200


def arr(shape=(4, 300), seed=0, nan_frac=0.15):
    a = np.random.RandomState(seed).rand(*shape)
    return np.where(a > nan_frac, a, np.nan)


def test_out_on_every_family(nb):
    a = arr()
    b = a**2 + 1
    # move: float64 loop, result cast into a float32 `out` (same_kind), the SAME object is returned
    o32 = np.empty(a.shape, np.float32)
    r = nb.move_mean(a, window=5, min_count=1, out=o32)
    assert r is o32
    np.testing.assert_allclose(o32, oracle.move_mean(a, window=5, min_count=1).astype(np.float32), rtol=1e-6, equal_nan=True)
    o = np.empty(a.shape)
    assert nb.move_corr(a, b, window=20, min_count=5, out=(o,)) is o  # 1-tuple form
    np.testing.assert_allclose(o, oracle.move_corr(a, b, window=20, min_count=5), rtol=1e-9, atol=1e-9, equal_nan=True)
    assert nb.move_exp_nanmean(a, alpha=0.3, out=o) is o
    np.testing.assert_allclose(o, oracle.move_exp_nanmean(a, alpha=0.3), rtol=1e-12, equal_nan=True)
    assert nb.ffill(a, limit=2, out=o) is o
    np.testing.assert_array_equal(o, oracle.ffill(a, limit=2))
    assert nb.bfill(a, out=o) is o
    np.testing.assert_array_equal(o, oracle.bfill(a))
    # matrix families
    ov = np.ascontiguousarray(a[:, :40].T)  # (obs, vars)
    m = np.empty((40, 4, 4))
    assert nb.move_covmatrix(ov, window=10, min_count=3, out=m) is m
    np.testing.assert_array_equal(m, oracle.move_covmatrix(ov, window=10, min_count=3))
    m2 = np.empty((4, 4))
    assert nb.nancorrmatrix(a, out=m2) is m2
    np.testing.assert_array_equal(m2, oracle.nancorrmatrix(a))
    m3 = np.empty((40, 4, 4))
    assert nb.move_exp_nancovmatrix(ov, alpha=0.4, out=m3) is m3
    np.testing.assert_array_equal(m3, oracle.move_exp_nancovmatrix(ov, alpha=0.4))
    # quantile: `out` has the gufunc's own layout (..., m), before the quantile axis moves first
    q = np.empty((4, 2))
    res = nb.nanquantile(a, [0.25, 0.5], axis=-1, out=q)
    np.testing.assert_array_equal(q.T, res)
    np.testing.assert_array_equal(res, oracle.nanquantile(a, [0.25, 0.5], axis=-1))
    # tensors in, tensor out
    import torch

    t = torch.from_numpy(a).cuda()
    to = torch.empty_like(t)
    assert nb.move_sum(t, window=4, min_count=1, out=to) is to
    np.testing.assert_allclose(to.cpu().numpy(), oracle.move_sum(a, window=4, min_count=1), rtol=1e-12, equal_nan=True)


def test_out_errors(nb):
    a = arr()
    with pytest.raises(TypeError, match="Cannot cast ufunc 'move_mean' output"):  # UFuncTypeError in NumPy
        nb.move_mean(a, window=5, out=np.empty(a.shape, np.int32))
    with pytest.raises(ValueError):
        nb.move_mean(a, window=5, out=np.empty((3, 300)))
    with pytest.raises(TypeError, match="unexpected keyword argument 'where'"):
        nb.move_mean(a, window=5, where=True)
    with pytest.raises(TypeError, match="unexpected keyword argument 'bogus'"):
        nb.ffill(a, bogus=1)
    nb.move_mean(a, window=5, order="F", subok=False)  # accepted by NumPy gufuncs, no effect


def test_dtype_and_casting(nb):
    a = arr()
    # exp family: every operand is a float, so dtype= picks the loop
    r = nb.move_exp_nanmean(a, alpha=0.5, dtype=np.float32)
    assert r.dtype == np.float32
    np.testing.assert_allclose(r, oracle.move_exp_nanmean(a.astype(np.float32), alpha=np.float32(0.5)), rtol=1e-5, equal_nan=True)
    assert nb.move_exp_nansum(a.astype(np.float32), alpha=np.float32(0.5), dtype=np.float64).dtype == np.float64
    # move / fill gufuncs also have int64 operands: only the float64 signature resolves
    assert nb.move_mean(a, window=3, min_count=1, dtype=np.float64).dtype == np.float64
    assert nb.move_mean(a.astype(np.float32), window=3, min_count=1, dtype=np.float64).dtype == np.float64
    with pytest.raises(TypeError, match="No loop matching the specified signature"):
        nb.move_mean(a, window=3, dtype=np.float32)
    with pytest.raises(TypeError, match="No loop matching the specified signature"):
        nb.ffill(a, dtype=np.float32)
    with pytest.raises(TypeError, match="No loop matching the specified signature"):
        nb.move_covmatrix(np.ascontiguousarray(a.T), window=3, dtype=np.float32)
    # casting= constrains the INPUT casts
    ai = np.arange(40).reshape(4, 10)
    assert nb.move_mean(ai, window=3, min_count=1, casting="safe").dtype == np.float64
    with pytest.raises(TypeError, match="Cannot cast ufunc 'move_mean' input 0"):
        nb.move_mean(ai, window=3, min_count=1, casting="no")
    assert nb.move_mean(a, window=3, min_count=1, casting="no").dtype == np.float64
    assert nb.move_mean(a.astype(np.float32), window=3, min_count=1, casting="no").dtype == np.float32
    with pytest.raises(ValueError):
        nb.move_mean(a, window=3, casting="bogus")


def test_big_endian_and_narrow_dtypes(nb):
    a = arr()
    be = a.astype(">f8")
    r = nb.move_mean(be, window=7, min_count=2)
    assert r.dtype == np.float64 and r.dtype.isnative
    np.testing.assert_allclose(r, oracle.move_mean(a, window=7, min_count=2), rtol=1e-12, equal_nan=True)
    np.testing.assert_array_equal(nb.ffill(a.astype(">f4")), oracle.ffill(a.astype(np.float32)))
    np.testing.assert_allclose(nb.group_nansum(be, np.arange(300) % 5, axis=-1), oracle.group_nansum(a, np.arange(300) % 5, axis=-1), rtol=1e-12)
    np.testing.assert_allclose(nb.nansum(be, axis=-1), oracle.nansum(a, axis=-1), rtol=1e-12)
    # loop selection over (float32, float64): the first loop the input casts to safely
    for dt, loop in ((np.int8, np.float32), (np.int16, np.float32), (np.uint8, np.float32), (np.uint16, np.float32),
                     (np.bool_, np.float32), (np.float16, np.float32), (np.int32, np.float64), (np.uint32, np.float64),
                     (np.int64, np.float64)):
        x = (np.random.RandomState(1).rand(3, 50) * 3).astype(dt)
        assert nb.move_mean(x, window=4, min_count=1).dtype == loop, dt
        np.testing.assert_allclose(nb.move_mean(x, window=4, min_count=1), oracle.move_mean(x.astype(loop), window=4, min_count=1), rtol=1e-6)
        assert nb.nanmean(x).dtype == loop, dt
        assert nb.nancovmatrix(x).dtype == loop, dt
        assert nb.move_exp_nanmean(x, alpha=0.5).dtype == np.float64  # a Python-float alpha forces the float64 loop
    # nanvar / nanstd carry an integer ddof operand: bool resolves to the float64 loop there
    xb = np.random.RandomState(2).rand(3, 50) > 0.5
    assert nb.nanvar(xb).dtype == np.float64 and nb.nanstd(xb).dtype == np.float64
    assert nb.nanvar(xb.astype(np.int8)).dtype == np.float32
    assert nb.nansum(xb.astype(np.int8)).dtype == np.int32 and nb.nanmax(xb.astype(np.int8)).dtype == np.int64


@pytest.mark.parametrize("shape", [(0, 5), (5, 0), (0,), (3, 0, 4)])
def test_empty_inputs(nb, shape):
    # zero-size inputs return an empty result like the reference
    a = np.empty(shape)
    for f in (nb.ffill, nb.bfill):
        r = f(a, axis=-1) if a.shape[-1] or a.ndim == 1 else f(a, axis=0)
        assert r.shape == a.shape and r.dtype == a.dtype
    r = nb.move_exp_nanmean(a, alpha=0.5)
    assert r.shape == a.shape
    r = nb.move_exp_nancorr(a, a, alpha=0.5)
    assert r.shape == a.shape
