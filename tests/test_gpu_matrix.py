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
