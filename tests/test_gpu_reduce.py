"""Plain NaN reductions on the GPU (nbg_reduce behind numbagg_b200's ndaggregate / ndreduce
mirrors) against the reference's frozen outputs (tests/golden/reduce.npz) and the oracle.

Parity classes (north_star): allnan, anynan, nancount, nanargmax, nanargmin, nanmax, nanmin
and every integer result are BIT-EXACT; nansum, nanmean, nanvar, nanstd of floats agree to
rtol 1e-12 (float64) / 1e-5 (float32).  Sums carry an absolute floor of rtol * sum(|x|) (a
cancelling sum cannot be relatively accurate in any summation order); float32 nansum gets the
reference's own rounding bound on top, because numbagg accumulates it sequentially IN
float32 (funcs.py:80, `asum = a.dtype.type(0)`) while the kernels accumulate in double."""

import numpy as np
import pytest

from oracle import oracle
from tests._golden import all_cases

pytestmark = pytest.mark.gpu

CASES = all_cases("reduce")
EXACT = {"allnan", "anynan", "nancount", "nanargmax", "nanargmin", "nanmax", "nanmin"}
RTOL = {np.dtype(np.float64): 1e-12, np.dtype(np.float32): 1e-5}
F32_EPS = float(np.finfo(np.float32).eps)


def check(func, a, got, exp, kwargs):
    got = np.asarray(got)
    exp = np.asarray(exp)
    assert got.shape == exp.shape, (func, got.shape, exp.shape)
    assert got.dtype == exp.dtype, (func, got.dtype, exp.dtype)
    if func in EXACT or exp.dtype.kind != "f":
        np.testing.assert_array_equal(got, exp)
        if exp.dtype.kind == "f":  # same zero sign is not required (max(-0.0, 0.0)), same NaN mask is
            assert np.array_equal(np.isnan(got), np.isnan(exp))
        return
    assert np.array_equal(np.isnan(got), np.isnan(exp)), f"{func}: NaN masks differ"
    rtol = RTOL[exp.dtype]
    axis = kwargs.get("axis")
    af = np.asarray(a, dtype=np.float64)
    with np.errstate(all="ignore"):
        if func in ("nansum", "nanmean"):
            mag = np.nansum(np.abs(np.where(np.isfinite(af), af, 0.0)), axis=axis)
            cnt = np.maximum(np.sum(~np.isnan(af), axis=axis), 1)
            atol = rtol * mag
            if func == "nansum" and exp.dtype == np.float32:
                atol = atol + cnt * F32_EPS * mag  # the reference's sequential float32 sum
            if func == "nanmean":
                atol = atol / cnt
        else:
            fin = np.where(np.isfinite(af), af, 0.0)
            atol = 1e-3 * rtol * float(np.max(np.abs(fin), initial=0.0)) ** 2
            if func == "nanstd":
                # compare as variances: the floor belongs to the variance
                got, exp, rtol = got.astype(np.float64) ** 2, exp.astype(np.float64) ** 2, 2 * rtol
    finite = np.isfinite(exp)
    assert np.array_equal(got[~finite], exp[~finite], equal_nan=True)
    err = np.abs(np.where(finite, got, 0) - np.where(finite, exp, 0))
    bound = atol + rtol * np.abs(np.where(finite, exp, 0))
    assert np.all(err <= bound), f"{func}: max excess {np.max(err - bound)} at {np.argmax(err - bound)}"


@pytest.mark.parametrize("case", CASES, ids=[c.id for c in CASES])
def test_cuda_matches_reference(case):
    import numbagg_b200 as nb

    got = getattr(nb, case.func)(*case.args, **case.kwargs)
    check(case.func, case.args[0], got, case.expected, case.kwargs)


def _data(shape, dtype, seed=0, nan_frac=0.15):
    rs = np.random.RandomState(seed)
    if np.dtype(dtype).kind == "f":
        a = rs.standard_normal(shape)
        a[rs.rand(*shape) < nan_frac] = np.nan
        return a.astype(dtype)
    return rs.randint(-1000, 1000, size=shape).astype(dtype)


FUNCS = oracle.AGGREGATION_FUNCS
FLOAT_ONLY = {"nanmean", "nanvar", "nanstd"}

# every kernel and geometry of nbg_reduce.cu: rows_tile and group (G = 1..32), stream (rows,
# narrow and wide columns, partial tiles, unaligned tails), rows_cta / cols for unaligned
# shapes, with and without segments, thread-per-output
SHAPES = [
    ((1000, 3), -1), ((500, 17), -1), ((300, 100), -1), ((64, 1000), -1), ((40, 4096), -1),
    ((700, 4099), -1), ((3, 70001), -1), ((1, 300007), -1), ((300007,), None),
    ((5000, 6), 0), ((20000, 3), 0), ((3000, 300), 0), ((9, 257, 130), 1), ((40, 12, 33), 1),
    ((50000, 4, 3), 1), ((2, 100000, 2), 1), ((70, 70, 70), (0, 1)), ((70, 70, 70), (1, 2)),
    ((30, 40, 50), (0, 2)), ((30, 40, 50), None),
    ((5, 40000), -1), ((33, 1024, 512), 1), ((4, 999, 64), 1), ((2, 33333, 8), 1), ((1, 50001, 4), 1),
]


@pytest.mark.parametrize("func", FUNCS)
@pytest.mark.parametrize("dtype", [np.float64, np.float32, np.int32, np.int64])
@pytest.mark.parametrize("shape,axis", SHAPES, ids=[f"{s}-{a}" for s, a in SHAPES])
def test_geometries_against_oracle(func, dtype, shape, axis):
    import numbagg_b200 as nb

    if np.dtype(dtype).kind != "f" and func in FLOAT_ONLY:
        pytest.skip("float-only loop")
    nan_frac = 0.15 if func not in ("nanargmax", "nanargmin") or min(shape) > 8 else 0.0
    a = _data(shape, dtype, seed=len(shape) * 7 + shape[0] % 13, nan_frac=nan_frac)
    kwargs = dict(axis=axis)
    try:
        exp = getattr(oracle, func)(a, **kwargs)
    except ValueError as e:
        with pytest.raises(ValueError, match=str(e)):
            getattr(nb, func)(a, **kwargs)
        return
    got = getattr(nb, func)(a, **kwargs)
    check(func, a, got, exp, kwargs)


@pytest.mark.parametrize("func", FUNCS)
def test_tensor_in_tensor_out_and_layouts(func):
    import torch

    import numbagg_b200 as nb

    a = _data((24, 36, 48), np.float64, seed=3)
    t = torch.from_numpy(a).cuda()
    views = {
        "c": (a, t),
        "f": (np.asfortranarray(a), t.permute(2, 1, 0).contiguous().permute(2, 1, 0)),
        "sliced": (a[:, ::2, 1:], t[:, ::2, 1:]),
    }
    for name, (an, tn) in views.items():
        for axis in (None, 0, 1, -1, (0, 1), (2, 1), (0, 2)):
            exp = getattr(oracle, func)(an, axis=axis)
            got = getattr(nb, func)(tn, axis=axis)
            assert isinstance(got, torch.Tensor) and got.is_cuda
            g = got.cpu().numpy()
            check(func, an, g[()] if g.ndim == 0 else g, exp, dict(axis=axis))


@pytest.mark.parametrize("func", ["nanvar", "nanstd"])
@pytest.mark.parametrize("ddof", [0, 1, 3, 50])
def test_ddof(func, ddof):
    import numbagg_b200 as nb

    a = _data((30, 40), np.float64, seed=9, nan_frac=0.5)
    check(func, a, getattr(nb, func)(a, axis=-1, ddof=ddof), getattr(oracle, func)(a, axis=-1, ddof=ddof), dict(axis=-1))


def test_variance_is_robust_where_a_one_pass_sum_of_squares_is_not():
    """funcs.py:115-117 keeps two passes for stability; the one-read kernel must not give
    that up: tiny spread on a huge offset, and a huge outlier first."""
    import numbagg_b200 as nb

    rs = np.random.RandomState(0)
    a = 1e9 + rs.standard_normal((4, 200001))
    a[1, 0] = 1e15
    a[2, ::3] = np.nan
    for func in ("nanvar", "nanstd"):
        check(func, a, getattr(nb, func)(a, axis=-1), getattr(oracle, func)(a, axis=-1), dict(axis=-1))
    got = nb.nanvar(a, axis=-1)
    assert abs(got[0] - 1.0) < 0.02 and abs(got[2] - 1.0) < 0.02


def test_errors_match_reference():
    import numbagg_b200 as nb

    a = _data((5, 40), np.float64, seed=1)
    a[3] = np.nan
    for f in ("nanargmax", "nanargmin"):
        with pytest.raises(ValueError, match="All-NaN slice encountered"):
            getattr(nb, f)(a, axis=-1)
        with pytest.raises(ValueError, match="All-NaN slice encountered"):
            getattr(nb, f)(np.full(7, np.nan))
    assert np.isnan(nb.nanmax(a, axis=-1)[3]) and np.isnan(nb.nanmin(a, axis=-1)[3])
    assert nb.nanmax(np.empty((0, 3)), axis=1).shape == (0,)
    assert nb.nansum(np.empty((0, 3)), axis=0).tolist() == [0.0, 0.0, 0.0]
    assert nb.allnan(np.empty((3, 0)), axis=1).tolist() == [True, True, True]
    assert np.isnan(nb.nanmean(np.empty((0,))))
    res = nb.nancount(np.empty((4, 0, 2)), axis=1)
    assert res.shape == (4, 2) and res.dtype == np.int64 and not res.any()


def test_first_extreme_wins_and_integers_compare_as_float64():
    import numbagg_b200 as nb

    a = np.zeros((3, 100000))
    a[0, [70000, 5, 99999]] = 7.0
    a[1, [12345, 54321]] = -2.0
    a[2, :] = np.nan
    a[2, 777] = -np.inf
    assert nb.nanargmax(a[:2], axis=-1).tolist() == [5, 0]
    assert nb.nanargmin(a[:2], axis=-1).tolist() == [0, 12345]
    assert nb.nanargmax(a[2]) == 777 and nb.nanargmin(a[2]) == 777
    big = np.array([2**53, 2**53 + 1, 2**53 - 1, 2**53 + 1], dtype=np.int64)  # float64 ties
    for f in ("nanargmax", "nanargmin", "nanmax", "nanmin"):
        assert getattr(nb, f)(big) == getattr(oracle, f)(big)


def test_shard_protocol_matches_single_pass():
    """nbg_reduce_partial on element shards + nbg_reduce_merge == nbg_reduce on the whole."""
    import torch

    from numbagg_b200 import decorators as dec

    a = _data((6, 50000), np.float64, seed=4)
    t = torch.from_numpy(a).cuda()
    cuts = [0, 1234, 20000, 20001, 50000]
    for func in FUNCS:
        whole = dec.run_reduce(func, t, (1,)).cpu().numpy()
        parts = []
        for lo, hi in zip(cuts[:-1], cuts[1:]):
            st, _ = dec.run_reduce_partial(func, t[:, lo:hi].contiguous(), (1,), index_offset=lo)
            parts.append(st)
        merged = dec.run_reduce_merge(func, np.dtype(np.float64), torch.stack(parts), a.shape[1]).cpu().numpy()
        if func in EXACT:
            np.testing.assert_array_equal(merged, whole)
        else:
            np.testing.assert_allclose(merged, whole, rtol=1e-12, equal_nan=True)


def test_full_size_properties():
    """BASELINE-sized input (10^7 elements): size-independent properties instead of the
    oracle -- counts add up, a permutation leaves order-free results unchanged, linearity."""
    import torch

    import numbagg_b200 as nb

    rs = np.random.RandomState(2)
    a = rs.standard_normal((1000, 10000))
    a[rs.rand(1000, 10000) < 0.1] = np.nan
    t = torch.from_numpy(a).cuda()
    cnt = nb.nancount(t, axis=-1)
    assert int(cnt.sum()) == int(np.count_nonzero(~np.isnan(a)))
    assert int(nb.nancount(t)) == int(cnt.sum())
    perm = torch.from_numpy(rs.permutation(10000)).cuda()
    tp = t[:, perm].contiguous()
    for f in ("nanmax", "nanmin", "nancount", "allnan", "anynan"):
        assert torch.equal(getattr(nb, f)(t, axis=-1), getattr(nb, f)(tp, axis=-1))
    torch.testing.assert_close(nb.nansum(tp, axis=-1), nb.nansum(t, axis=-1), rtol=1e-12, atol=1e-10)
    torch.testing.assert_close(nb.nanvar(tp, axis=-1), nb.nanvar(t, axis=-1), rtol=1e-12, atol=0)
    torch.testing.assert_close(nb.nansum(t * 2.0, axis=0), nb.nansum(t, axis=0) * 2.0, rtol=1e-13, atol=0)
    am = nb.nanargmax(t, axis=-1)
    assert torch.equal(t.gather(1, am[:, None])[:, 0], nb.nanmax(t, axis=-1))
    assert torch.equal(nb.nanmax(t), nb.nanmax(nb.nanmax(t, axis=0)))


def test_billion_element_consistency():
    """10^9 float32 elements (BASELINE config-3 scale): the one-row segmented path, the
    many-rows path and the column path must agree with each other and with the construction."""
    import torch

    import numbagg_b200 as nb

    g = torch.Generator(device="cuda").manual_seed(7)
    t = torch.rand((1000, 1_000_000), generator=g, device="cuda", dtype=torch.float32)
    t[t < 0.05] = float("nan")
    n_valid = int(torch.count_nonzero(~torch.isnan(t)))
    assert int(nb.nancount(t)) == n_valid
    assert int(nb.nancount(t, axis=-1).sum()) == n_valid and int(nb.nancount(t, axis=0).sum()) == n_valid
    total = float(nb.nansum(t.view(-1)))
    by_rows = float(nb.nansum(t, axis=-1).double().sum())
    by_cols = float(nb.nansum(t, axis=0).double().sum())
    assert abs(total - by_rows) <= 1e-6 * abs(by_rows) and abs(by_cols - by_rows) <= 1e-6 * abs(by_rows)
    flat_arg = int(nb.nanargmax(t.view(-1)))
    assert float(t.view(-1)[flat_arg]) == float(nb.nanmax(t))
    assert flat_arg == int(torch.nonzero(t.view(-1) == nb.nanmax(t))[0])  # the FIRST maximum
    v_all = float(nb.nanvar(t.view(-1)))
    assert abs(v_all - 1.0 / 12.0 * (0.95**2) - 0.0) < 0.01  # uniform(0.05, 1): var = 0.95^2 / 12


@pytest.mark.parametrize("seed", range(6))
def test_random_shapes_axes_and_layouts(seed):
    """Randomised sweep over the geometry decisions of nbg_reduce.cu (kernel choice, segment
    counts, alignment fallbacks, tile widths at their thresholds)."""
    import numbagg_b200 as nb

    rs = np.random.RandomState(1000 + seed)
    edge = [1, 2, 3, 4, 7, 8, 16, 17, 31, 32, 33, 63, 64, 65, 100, 255, 256, 257, 511, 512, 1000, 1023, 1024, 1025,
            2047, 2048, 4095, 4096, 4097, 8191, 16384, 16385, 40000]
    for _ in range(40):
        nd = rs.randint(1, 4)
        budget = 400_000
        shape = []
        for _d in range(nd):
            s = int(edge[rs.randint(len(edge))])
            s = max(1, min(s, budget))
            shape.append(s)
            budget = max(1, budget // s)
        shape = tuple(shape)
        dtype = [np.float64, np.float32, np.int32, np.int64][rs.randint(4)]
        a = _data(shape, dtype, seed=rs.randint(1 << 30), nan_frac=[0.0, 0.1, 0.6][rs.randint(3)])
        if rs.rand() < 0.3 and nd > 1:
            perm = rs.permutation(nd)
            a = np.ascontiguousarray(a.transpose(perm)).transpose(np.argsort(perm))
        if rs.rand() < 0.2 and a.shape[-1] > 2:
            a = a[..., 1:]  # unaligned, non-contiguous view
        k = rs.randint(0, nd + 1)
        axis = None if k == 0 else tuple(int(x) for x in rs.choice(nd, size=k, replace=False))
        if axis is not None and len(axis) == 1:
            axis = axis[0]
        for func in FUNCS:
            if np.dtype(dtype).kind != "f" and func in FLOAT_ONLY:
                continue
            try:
                exp = getattr(oracle, func)(a, axis=axis)
            except ValueError as e:
                with pytest.raises(ValueError, match=str(e)[:20]):
                    getattr(nb, func)(a, axis=axis)
                continue
            got = getattr(nb, func)(a, axis=axis)
            try:
                check(func, a, got, exp, dict(axis=axis))
            except AssertionError as err:
                raise AssertionError(f"{func} shape={a.shape} strides={a.strides} dtype={a.dtype} axis={axis}: {err}")
