"""CPU-side checks for the plain-reduction row: the host layout logic
(ReduceView: which (outer, n, inner) view -- or copy -- represents `axis`), dtype/loop
selection and the validation errors the reference raises before any loop runs.  The oracle
itself is pinned against the reference's outputs in tests/test_oracle_golden.py."""

import itertools

import numpy as np
import pytest
import torch

from numbagg_b200 import decorators as dec
from oracle import oracle


def _layouts(shape):
    nd = len(shape)
    for perm in itertools.permutations(range(nd)):
        yield perm


@pytest.mark.parametrize("shape", [(5,), (3, 4), (2, 3, 4), (2, 1, 3, 2)])
def test_reduce_view_matches_moveaxis(shape):
    """For every memory layout and every ordered choice of axes, element k of the middle axis
    of the view is element k of numpy's moveaxis+flatten, and restore() puts outputs back in
    the batch order."""
    nd = len(shape)
    base = np.arange(int(np.prod(shape)), dtype=np.float64).reshape(shape)
    for perm in _layouts(shape):
        a = np.ascontiguousarray(base.transpose(perm)).transpose(np.argsort(perm))
        assert a.shape == tuple(shape)
        t = torch.from_numpy(a)
        for r in range(1, nd + 1):
            for axes in itertools.permutations(range(nd), r):
                view = dec.ReduceView(t, axes)
                assert view.t.is_contiguous()
                cube = view.t.reshape(view.outer, view.n, view.inner)
                moved = np.moveaxis(a, axes, range(nd - r, nd))
                bshape = moved.shape[: nd - r]
                want = moved.reshape(bshape + (-1,))
                assert view.n == want.shape[-1]
                for k in range(view.n):
                    got = view.restore(cube[:, k, :].reshape(-1)).numpy()
                    np.testing.assert_array_equal(got, want[..., k])


def test_reduce_view_copies_only_when_needed():
    t = torch.zeros((4, 5, 6))
    for axes in [(0,), (1,), (2,), (0, 1), (1, 2), (0, 1, 2)]:
        assert dec.ReduceView(t, axes).t.data_ptr() == t.data_ptr()
    assert dec.ReduceView(t, (0, 2)).t.data_ptr() != t.data_ptr()   # not adjacent in memory
    assert dec.ReduceView(t, (2, 1)).t.data_ptr() != t.data_ptr()   # adjacent, but reversed order
    f = t.permute(2, 1, 0)  # F-ordered view
    assert dec.ReduceView(f, (1, 0)).t.data_ptr() == t.data_ptr()
    sliced = t[:, ::2]
    assert dec.ReduceView(sliced, (1,)).t.data_ptr() != t.data_ptr()


def test_reduce_loop_dtype_follows_numpy_casting():
    f = dec._reduce_loop_dtype
    assert f("nansum", np.dtype("f2")) == np.float32
    assert f("nansum", np.dtype("f4")) == np.float32
    assert f("nansum", np.dtype("i1")) == np.int32
    assert f("nansum", np.dtype("u2")) == np.int32
    assert f("nansum", np.dtype("u4")) == np.int64
    assert f("nansum", np.dtype("i8")) == np.int64
    assert f("nansum", np.dtype("?")) == np.int32
    assert f("nanmean", np.dtype("i4")) == np.float64
    assert f("nanstd", np.dtype("f4")) == np.float32
    with pytest.raises(TypeError):
        f("nansum", np.dtype("c16"))


def test_oracle_reduce_errors_match_reference_messages():
    with pytest.raises(ValueError, match="All-NaN slice encountered"):
        oracle.nanargmax(np.array([np.nan, np.nan]))
    with pytest.raises(ValueError, match="All-NaN slice encountered"):
        oracle.nanargmin(np.empty((3, 0)), axis=-1)
    with pytest.raises(ValueError, match="fmax which has no identity"):
        oracle.nanmax(np.empty((0,)))
    with pytest.raises(ValueError, match="fmin which has no identity"):
        oracle.nanmin(np.empty((2, 0)), axis=1)
    assert oracle.nanmax(np.empty((0, 3)), axis=1).shape == (0,)


def test_host_validation_needs_no_gpu():
    import numbagg_b200 as nb

    with pytest.raises(TypeError, match="must be arrays"):
        nb.nansum([1.0, 2.0])
    with pytest.raises(np.exceptions.AxisError):
        nb.nanmax(np.zeros((2, 3)), axis=2)
    with pytest.raises(ValueError, match="repeated axis"):
        nb.nanmin(np.zeros((2, 3)), axis=(0, 0))
    with pytest.raises(ValueError, match="fmax which has no identity"):
        nb.nanmax(np.empty((3, 0)), axis=-1)
    with pytest.raises(ValueError, match="All-NaN slice encountered"):
        nb.nanargmax(np.empty((0,)))
    assert "ddof" not in nb.nansum.__signature__.parameters
    assert "ddof" in nb.nanvar.__signature__.parameters
