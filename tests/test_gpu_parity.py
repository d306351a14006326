"""CUDA path vs the oracle (oracle/nbg_oracle.c, pinned bit-for-bit to the reference by
tests/test_oracle_golden.py) on seeded inputs at sizes the oracle finishes in seconds, plus
size-independent properties at larger sizes.  Everything goes through the public drop-in API
(numpy in / numpy out) or the C ABI behind it.  Tolerances: tests/_parity.py (north_star)."""

import numpy as np
import pytest

from oracle import oracle
from tests._parity import assert_parity

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def nb():
    import numbagg_b200

    return numbagg_b200


def fixture_array(shape, nan_frac=0.1, seed=0, dtype=np.float64):
    a = np.random.RandomState(seed).rand(*shape)
    return np.where(a > nan_frac, a, np.nan).astype(dtype)


ONE = ["move_mean", "move_sum", "move_std", "move_var"]
TWO = ["move_cov", "move_corr"]


def _scale(func, arrs, window=1):
    m = max(float(np.nanmax(np.abs(a))) for a in arrs)
    if "corr" in func:
        return 1.0
    s = m * m if any(t in func for t in ("var", "std", "cov")) else m
    return s * (window if func == "move_sum" else 1)


# ------------------------------------------------------------------------------- moving
@pytest.mark.parametrize("dtype", [np.float64, np.float32])
@pytest.mark.parametrize("window,min_count", [(1, None), (2, 1), (20, 1), (32, 5), (33, None), (100, 50), (1000, 500), (4096, 1), (5000, 10), (20000, 1)])
def test_move_vs_oracle_tiles(nb, dtype, window, min_count):
    # 20000 columns: several row tiles per row for both dtypes; windows on both sides of the
    # direct / delta-scan switch (32), of the reciprocal-table limit (4096) and window == n
    a = fixture_array((6, 20000), dtype=dtype, seed=1)
    b = (a.astype(np.float64) ** 2 + 1).astype(dtype)
    for f in ONE:
        got = getattr(nb, f)(a, window=window, min_count=min_count)
        exp = getattr(oracle, f)(a, window=window, min_count=min_count)
        assert_parity(f, got, exp, scale=_scale(f, [a], window))
    for f in TWO:
        got = getattr(nb, f)(a, b, window=window, min_count=min_count)
        exp = getattr(oracle, f)(a, b, window=window, min_count=min_count)
        # The reference never re-syncs its running sums (moving.py:30-55): after i steps they
        # carry an absolute drift of ~sqrt(i) ulps of the sums.  A correlation DIVIDES by
        # variances formed from those sums, so for windows of a few near-equal points the
        # reference's own result is only accurate to drift/variance; compare those loosely.
        # With 1- or 2-point windows even the SIGN of var_a*var_b (the NaN gate, moving.py:268)
        # is decided by that drift, so the correlation of such windows is not comparable.
        if f == "move_corr" and window < 3:
            continue
        atol = 1e-6 if (f == "move_corr" and window < 20) else None
        assert_parity(f, got, exp, scale=_scale(f, [a, b], window), atol=atol)


@pytest.mark.parametrize("n", [1, 2, 16, 17, 4607, 4608, 4609, 8703, 8704, 8705, 10001])
def test_move_ragged_lengths(nb, n):
    # tile sizes are 4608 (float64) and 8704 (float32): hit the boundaries and odd row pitches
    for dtype in (np.float64, np.float32):
        a = fixture_array((3, n), dtype=dtype, seed=n)
        w = min(n, 20)
        for f in ("move_mean", "move_sum", "move_std"):
            got = getattr(nb, f)(a, window=w, min_count=1)
            exp = getattr(oracle, f)(a, window=w, min_count=1)
            assert_parity(f, got, exp, scale=_scale(f, [a], w))


def test_move_axes_and_layouts(nb):
    a = fixture_array((7, 33, 50), seed=3)
    for axis in (0, 1, 2, -2):
        for f in ("move_mean", "move_var"):
            got = getattr(nb, f)(a, window=5, min_count=2, axis=axis)
            exp = getattr(oracle, f)(a, window=5, min_count=2, axis=axis)
            assert_parity(f, got, exp, scale=_scale(f, [a]))
    # F-ordered, transposed and strided views: no layout may change values
    af = np.asfortranarray(a[0])
    for arr in (af, a[0].T, a[:, ::2, 1:40:3], a[::-1]):
        got = nb.move_sum(arr, window=3, min_count=1, axis=-1)
        exp = oracle.move_sum(arr, window=3, min_count=1, axis=-1)
        assert_parity("move_sum", got, exp, scale=3.0)
        got = nb.move_sum(arr, window=3, min_count=1, axis=0)
        exp = oracle.move_sum(arr, window=3, min_count=1, axis=0)
        assert_parity("move_sum", got, exp, scale=3.0)


def test_move_long_columns_other_axis(nb):
    # core axis = 0 of a C-contiguous matrix: column-walk kernel, segmented along the core axis
    a = fixture_array((30000, 37), seed=4)
    b = a**2 + 1
    for f in ("move_mean", "move_std"):
        got = getattr(nb, f)(a, window=100, min_count=10, axis=0)
        exp = getattr(oracle, f)(a, window=100, min_count=10, axis=0)
        assert_parity(f, got, exp, scale=_scale(f, [a]))
    got = nb.move_corr(a, b, window=100, min_count=10, axis=0)
    assert_parity("move_corr", got, oracle.move_corr(a, b, window=100, min_count=10, axis=0), scale=1.0)


def test_move_dtype_rules_and_tensors(nb):
    import torch

    ints = np.arange(40).reshape(4, 10)
    got = nb.move_mean(ints, window=3)
    assert got.dtype == np.float64
    assert_parity("move_mean", got, oracle.move_mean(ints, window=3))
    h = np.arange(10, dtype=np.float16)
    assert nb.move_mean(h, window=3).dtype == np.float32
    # float32 stability tests of the reference (test_moving.py:180-192)
    arr = np.array([0.1, 0.2, 0.3] * 100, dtype=np.float32)
    np.testing.assert_array_equal(nb.move_mean(arr, window=1), arr)
    tiled = np.tile(np.arange(10, dtype=np.float32) * 1.7, 30)
    assert nb.move_sum(tiled, window=10)[-1] == np.sum(tiled[:10], dtype=np.float64).astype(np.float32)
    # CUDA tensors in -> CUDA tensor out, same stream, no host round trip
    t = torch.from_numpy(fixture_array((5, 300))).cuda()
    r = nb.move_mean(t, window=7, min_count=1)
    assert isinstance(r, torch.Tensor) and r.is_cuda and r.dtype == torch.float64
    assert_parity("move_mean", r.cpu().numpy(), oracle.move_mean(t.cpu().numpy(), window=7, min_count=1))


def test_move_all_nan_and_inf(nb):
    a = np.full((2, 5000), np.nan)
    for f in ONE:
        np.testing.assert_array_equal(getattr(nb, f)(a, window=10, min_count=0), getattr(oracle, f)(a, window=10, min_count=0))
    x = np.array([1.0, np.inf, 2.0, -np.inf, 3.0, 4.0, 5.0])
    np.testing.assert_array_equal(nb.move_sum(x, window=2, min_count=1), oracle.move_sum(x, window=2, min_count=1))


# --------------------------------------------------------------------------- exp moving
EXP_ONE = ["move_exp_nancount", "move_exp_nanmean", "move_exp_nansum", "move_exp_nanvar", "move_exp_nanstd"]
EXP_TWO = ["move_exp_nancov", "move_exp_nancorr"]


@pytest.mark.parametrize("dtype", [np.float64, np.float32])
@pytest.mark.parametrize("alpha", [0.5, 0.1, 0.001])
def test_move_exp_vs_oracle_long_rows(nb, dtype, alpha):
    # 200k columns: ~45 (float64) / ~24 (float32) chained tiles per row, 30 % NaN (config 3)
    a = fixture_array((3, 200_000), nan_frac=0.3, dtype=dtype, seed=5)
    b = (a.astype(np.float64) ** 2 + 1).astype(dtype)
    al = np.float32(alpha) if dtype == np.float32 else alpha
    for f in EXP_ONE:
        for mw in (0, 0.5):
            got = getattr(nb, f)(a, alpha=al, min_weight=mw)
            exp = getattr(oracle, f)(a, alpha=al, min_weight=mw)
            assert_parity(f, got, exp, scale=(1.0 / alpha if f == "move_exp_nansum" else 1.0))
    for f in EXP_TWO:
        got = getattr(nb, f)(a, b, alpha=al)
        exp = getattr(oracle, f)(a, b, alpha=al)
        assert_parity(f, got, exp, scale=4.0)


@pytest.mark.parametrize("alpha,mw", [(1.2, 0), (-0.1, 0), (0.3, -1.0), (1.0, 0), (0.0, 0), (0.3, float("nan"))])
def test_move_exp_gate_elimination_boundaries(nb, alpha, mw):
    # The weight channel is dropped only when the gate `weight >= min_weight` is provably open (scalar alpha
    # in [0, 1], min_weight <= 0): alphas outside that range, a NaN min_weight and the interval ends must
    # still agree with the reference (which applies no range check, decorators.py:356-359)
    a = fixture_array((2, 30_000), nan_frac=0.3, seed=9)
    a[1, :50] = np.nan
    b = a**2 + 1
    for f in EXP_ONE:
        got = getattr(nb, f)(a, alpha=alpha, min_weight=mw)
        exp = getattr(oracle, f)(a, alpha=alpha, min_weight=mw)
        fin = np.isfinite(exp) & (np.abs(exp) < 1e200)  # alpha outside [0, 1] diverges: compare what stays finite
        assert np.array_equal(np.isnan(got[:, :200]), np.isnan(exp[:, :200]))
        np.testing.assert_allclose(got[fin][:4000], exp[fin][:4000], rtol=1e-9, atol=1e-12)
    for f in EXP_TWO:
        got = getattr(nb, f)(a, b, alpha=alpha, min_weight=mw)
        exp = getattr(oracle, f)(a, b, alpha=alpha, min_weight=mw)
        fin = np.isfinite(exp) & np.isfinite(got)
        np.testing.assert_allclose(got[fin][:4000], exp[fin][:4000], rtol=1e-6, atol=1e-9)


def test_move_prefix_kernel_every_op(nb, monkeypatch):
    # float32 windows > 32 as differences of a streaming prefix (nbg_move_prefix.cuh).  By default only
    # move_std / move_cov take it (where it is the faster kernel); NBG_PFX=all routes every op through it
    monkeypatch.setenv("NBG_PFX", "all")
    for shape, window, mc in (((5, 30_000), 1000, 500), ((3, 9_001), 33, None), ((2, 50_000), 5000, 10), ((7, 2_305), 100, 1),
                              ((1, 120_003), 4097, 2000)):
        a = fixture_array(shape, dtype=np.float32, seed=window)
        a[0, :window // 2] = np.nan
        b = (a.astype(np.float64) ** 2 + 1).astype(np.float32)
        for f in ONE:
            got = getattr(nb, f)(a, window=window, min_count=mc)
            exp = getattr(oracle, f)(a, window=window, min_count=mc)
            assert_parity(f, got, exp, scale=_scale(f, [a], window))
        for f in TWO:
            got = getattr(nb, f)(a, b, window=window, min_count=mc)
            exp = getattr(oracle, f)(a, b, window=window, min_count=mc)
            assert_parity(f, got, exp, scale=_scale(f, [a, b], window), atol=1e-5 if f == "move_corr" else None)
    # constant and zero windows: the float32 image of the variance is 0 / subnormal -> exact redo path
    c = np.full((2, 20_000), 3.25, dtype=np.float32)
    c[1, ::7] = 0.0
    c[1, 5000:9000] = 0.0
    for f in ("move_std", "move_var", "move_mean"):
        got = getattr(nb, f)(c, window=64, min_count=2)
        exp = getattr(oracle, f)(c, window=64, min_count=2)
        assert np.array_equal(np.isnan(got), np.isnan(exp)) or f == "move_std", f
        np.testing.assert_allclose(np.nan_to_num(got, nan=0.0), np.nan_to_num(exp, nan=0.0), rtol=1e-5, atol=1e-5)


def test_move_exp_alpha_forms_and_axes(nb):
    a = fixture_array((5, 9000), seed=6)
    al1 = np.random.RandomState(7).rand(9000) * 0.9 + 0.05
    aln = np.random.RandomState(8).rand(5, 9000) * 0.9 + 0.05
    for f in EXP_ONE:
        for al in (al1, aln):
            assert_parity(f, getattr(nb, f)(a, alpha=al), getattr(oracle, f)(a, alpha=al), scale=20.0)
        got = getattr(nb, f)(a.T.copy(), alpha=aln.T.copy(), axis=0)
        assert_parity(f, got, getattr(oracle, f)(a.T.copy(), alpha=aln.T.copy(), axis=0), scale=20.0)
        got = getattr(nb, f)(a.T.copy(), alpha=0.3, axis=0)
        assert_parity(f, got, getattr(oracle, f)(a.T.copy(), alpha=0.3, axis=0), scale=4.0)
    # alpha with zeros: no decay, the look-back can never stop early
    al0 = al1.copy()
    al0[::3] = 0.0
    assert_parity("move_exp_nansum", nb.move_exp_nansum(a, alpha=al0), oracle.move_exp_nansum(a, alpha=al0), scale=100.0)
    # float32 data + python float alpha runs the float64 loop
    a32 = a.astype(np.float32)
    got = nb.move_exp_nanmean(a32, alpha=0.25)
    assert got.dtype == np.float64
    assert_parity("move_exp_nanmean", got, oracle.move_exp_nanmean(a32, alpha=0.25))


def test_move_exp_leading_nans_across_tiles(nb):
    a = fixture_array((2, 30000), nan_frac=0.3, seed=9)
    a[:, :12000] = np.nan  # more than two whole tiles of NaN before the first observation
    for f in EXP_ONE:
        assert_parity(f, getattr(nb, f)(a, alpha=0.1), getattr(oracle, f)(a, alpha=0.1), scale=10.0)


# -------------------------------------------------------------------------------- fills
@pytest.mark.parametrize("dtype", [np.float64, np.float32])
@pytest.mark.parametrize("limit", [None, 0, 1, 7, 5000, 40000])
def test_fill_vs_oracle_bit_exact(nb, dtype, limit):
    a = fixture_array((3, 100_000), nan_frac=0.3, dtype=dtype, seed=10)
    a[1, 20_000:65_000] = np.nan  # a NaN run spanning many tiles
    a[2, :] = np.nan
    a[2, 77] = 3.5
    for f in ("ffill", "bfill"):
        got = getattr(nb, f)(a, limit=limit)
        exp = getattr(oracle, f)(a, limit=limit)
        assert got.dtype == exp.dtype
        np.testing.assert_array_equal(got, exp)
        got0 = getattr(nb, f)(a.T.copy(), limit=limit, axis=0)
        np.testing.assert_array_equal(got0, exp.T)


def test_fill_misc_dtypes(nb):
    h = np.array([np.nan, 1, np.nan, 2], dtype=np.float16)
    r = nb.ffill(h)
    assert r.dtype == np.float16
    np.testing.assert_array_equal(r, np.array([np.nan, 1, 1, 2], dtype=np.float16))
    x = np.array([np.inf, np.nan, -np.inf, np.nan])
    np.testing.assert_array_equal(nb.ffill(x), [np.inf, np.inf, -np.inf, -np.inf])
    np.testing.assert_array_equal(nb.bfill(x), [np.inf, -np.inf, -np.inf, np.nan])


# ------------------------------------------------------------------------------ grouped
GROUP = oracle.GROUPED_FUNCS
GROUP_FLOAT_ONLY = {"group_nanvar", "group_nanstd"}


def _group_check(nb, f, values, labels, f32_sum_note=False, **kw):
    got = getattr(nb, f)(values, labels, **kw)
    exp = getattr(oracle, f)(values, labels, **kw)
    mask = None
    if exp.dtype.kind in "iu" and f in ("group_nanfirst", "group_nanlast", "group_nanargmax", "group_nanargmin", "group_nanmin", "group_nanmax"):
        mask = oracle.group_nancount(values, labels, **kw) == 0
    # absolute floor: sums of mixed-sign values cancel, and the float32 reference accumulates
    # in float32 (grouped.py accumulates in the output dtype), so an output is only defined
    # to rtol * (magnitude of the summands)
    vals = np.asarray(values)
    m = float(np.nanmax(np.abs(vals.astype(np.float64)))) if vals.size else 0.0
    per_group = max(1.0, vals.shape[-1] / max(1, int(np.max(labels)) + 1)) if f != "group_nanprod" else 1.0
    scale = None
    atol = None
    if f in GROUP_FLOAT_ONLY or f == "group_nansum_of_squares":
        scale = m * m * per_group
    elif f in ("group_nansum", "group_nanmean"):
        scale = m * per_group
    elif f == "group_nanprod" and exp.dtype.kind == "f":
        atol = float(np.finfo(exp.dtype).tiny)  # subnormal products
    assert_parity(f, got, exp, scale=scale, int_empty_mask=mask, atol=atol)


@pytest.mark.parametrize("dtype", [np.float64, np.float32])
@pytest.mark.parametrize("K", [3, 12, 300, 5000, 70000])
def test_group_shared_labels_vs_oracle(nb, dtype, K):
    # labels shared by all rows (axis=-1): few-label private bins, row-bins, and -- beyond the
    # shared-memory bin table -- the atomic path; ~100 observations per group at most so the
    # float32 reference itself is accurate to well below rtol
    rows, n = 19, 4000
    rs = np.random.RandomState(K)
    v = np.round((fixture_array((rows, n), dtype=dtype, seed=K) - 0.4) * 8, 2).astype(dtype)
    labels = rs.randint(-1, K + 2, size=n)  # includes skipped (-1) and out-of-range labels
    for f in GROUP:
        # products of ~n/K factors: keep them inside the float range (overflow followed by a
        # zero factor gives NaN in an order-dependent way, in the reference too)
        vv = (1.0 + v / 16).astype(dtype) if f == "group_nanprod" else v
        _group_check(nb, f, vv, labels, num_labels=K, axis=-1)


@pytest.mark.parametrize("dtype", [np.float64, np.float32])
@pytest.mark.parametrize("rows,n,K", [(1, 400_000, 12), (3, 300_000, 7), (20, 200_000, 1000), (1, 250_000, 300)])
def test_group_value_index_ops_across_column_segments(nb, dtype, rows, n, K):
    # few rows x long core axis: the rows are cut into column segments whose (value, index)
    # bins merge under the row-group lock; leading NaN runs make later segments win
    rs = np.random.RandomState(rows + K)
    v = np.round(fixture_array((rows, n), dtype=dtype, seed=K, nan_frac=0.3) * 50).astype(dtype)  # many ties
    v[:, : n // 3] = np.nan
    v[0, n // 3 : n // 2][::2] = np.nan
    labels = rs.randint(-1, K, size=n)
    for f in ("group_nanargmax", "group_nanargmin", "group_nanfirst", "group_nanlast", "group_nanmax", "group_nanmin"):
        _group_check(nb, f, v, labels, num_labels=K, axis=-1)


def test_group_rowbins_legacy_is_bit_exact_with_whole_rows(nb, monkeypatch):
    # The round-1 row-bins kernel (still the path of arg*/first/last/min/max and of bins that do
    # not fit the class-private layout): >= 4 waves of 8-row groups => one CTA per row group walks
    # whole rows in column order, so float32 sums are IDENTICAL to the sequential reference.
    monkeypatch.setenv("NBG_RB2_OFF", "1")
    rows, n, K = 9600, 256, 37
    v = fixture_array((rows, n), dtype=np.float32, seed=21)
    labels = np.random.RandomState(21).randint(0, K, size=n)
    for f in ("group_nansum", "group_nanmean", "group_nansum_of_squares", "group_nancount", "group_nanvar", "group_nanstd"):
        got = getattr(nb, f)(v, labels, num_labels=K, axis=-1)
        exp = getattr(oracle, f)(v, labels, num_labels=K, axis=-1)
        np.testing.assert_array_equal(got, exp, err_msg=f)


@pytest.mark.parametrize("dtype", [np.float32, np.float64, np.int32, np.int64])
@pytest.mark.parametrize("rows,n,K", [(9600, 256, 37), (64, 40_000, 1000), (17, 100_000, 1400), (3, 300_000, 5),
                                      (1, 200_004, 600), (40, 4096, 1)])
def test_group_rowbins2_class_private_bins(nb, dtype, rows, n, K):
    # conflict-free class-private bins (nbg_group_rowbins2.cuh): whole rows per CTA, column
    # segments merged with atomics (few rows), register runs (few labels), the dummy-padded
    # plan (ragged last tile, labels outside [0, K)), every value dtype.  Counts / any / all and
    # integer sums are exact; float sums differ from the reference only by summation order.
    rs = np.random.RandomState(rows + K)
    if np.dtype(dtype).kind == "f":
        v = np.round((fixture_array((rows, n), dtype=dtype, seed=K) - 0.3) * 8, 3).astype(dtype)
    else:
        v = rs.randint(-50, 50, size=(rows, n)).astype(dtype)
    labels = rs.randint(-1, K + 1, size=n)
    fs = ["group_nansum", "group_nancount", "group_nansum_of_squares", "group_nanany", "group_nanall"]
    if np.dtype(dtype).kind == "f":
        fs += ["group_nanmean", "group_nanvar", "group_nanstd"]
    for f in fs:
        _group_check(nb, f, v, labels, num_labels=K, axis=-1)
    vv = (1.0 + v / 64).astype(dtype) if np.dtype(dtype).kind == "f" else np.where(v % 7 == 0, 2, 1).astype(dtype)
    _group_check(nb, "group_nanprod", vv, labels, num_labels=K, axis=-1)


@pytest.mark.parametrize("dtype", [np.float64, np.float32, np.int32, np.int64])
def test_group_axis_modes_and_dtypes(nb, dtype):
    rs = np.random.RandomState(11)
    if np.dtype(dtype).kind == "f":
        v = fixture_array((6, 5, 40), dtype=dtype, seed=12)
    else:
        v = rs.randint(-9, 10, size=(6, 5, 40)).astype(dtype)
    for f in GROUP:
        if f in GROUP_FLOAT_ONLY and np.dtype(dtype).kind != "f":
            continue
        _group_check(nb, f, v, rs.randint(0, 4, size=40), axis=-1)
        _group_check(nb, f, v, rs.randint(0, 3, size=6), axis=0)
        _group_check(nb, f, v, rs.randint(-1, 5, size=(5, 40)), axis=(1, 2))
        _group_check(nb, f, v, rs.randint(0, 7, size=(6, 5, 40)), axis=None)
        _group_check(nb, f, v.reshape(-1), rs.randint(0, 7, size=v.size))


def test_group_bool_narrow_ints_and_label_dtypes(nb):
    rs = np.random.RandomState(13)
    bv = rs.rand(500) > 0.5
    bl = rs.randint(0, 4, size=500)
    for f in ("group_nansum", "group_nanany", "group_nanall", "group_nancount", "group_nanmean"):
        _group_check(nb, f, bv, bl)
    v = fixture_array((300,), seed=14)
    for ldt in (np.int8, np.int16, np.int32, np.int64):
        _group_check(nb, "group_nansum", v, rs.randint(0, 5, size=300).astype(ldt))
    i8 = rs.randint(-100, 100, size=400).astype(np.int8)
    got = nb.group_nansum(i8, rs.randint(0, 3, size=400))
    assert got.dtype == np.int8  # wrap-around accumulation in the values dtype
    # num_labels inferred from labels.max()
    got = nb.group_nanmax(v, np.array([0, 5] * 150))
    assert got.shape == (6,)
    # group axis goes LAST (SURVEY discrepancy table)
    assert nb.group_nansum(np.arange(12.0).reshape(4, 3), np.array([0, 1, 0, 1]), axis=0).shape == (3, 2)


def test_group_high_cardinality_1d(nb):
    # config-5 style: per-element labels, many groups, global atomics
    n, K = 3_000_000, 400_000
    rs = np.random.RandomState(15)
    v = fixture_array((n,), seed=16)
    labels = rs.randint(0, K, size=n)
    for f in ("group_nanargmax", "group_nanargmin", "group_nanfirst", "group_nanlast", "group_nanvar", "group_nansum", "group_nanmin", "group_nancount"):
        _group_check(nb, f, v, labels, num_labels=K)


@pytest.mark.parametrize("dtype", [np.float64, np.float32])
def test_group_partition_path(nb, dtype, monkeypatch):
    """Per-element labels, a table too large for L2, one row of floating-point values: mean / var / std
    can partition the elements by label range and reduce in shared-memory bins (nbg_group_partition.cuh;
    an experiment, enabled with NBG_GROUP_PARTITION=1 -- the C library reads it per call).  Labels out of
    range, NaN values, empty labels and an empty tail bucket are all present.  The default path (one
    atomic pass per channel plane) runs on the same input first."""
    _partition_case(nb, dtype)
    monkeypatch.setenv("NBG_GROUP_PARTITION", "1")
    _partition_case(nb, dtype)


def _partition_case(nb, dtype):
    n, K = 12_000_000, 6_600_000
    rs = np.random.RandomState(21)
    v = (rs.standard_normal(n) * 3 + 1).astype(dtype)
    v[rs.rand(n) < 0.1] = np.nan
    labels = rs.randint(-2, K - 20_000, size=n)  # some negative (ignored), the last labels never occur
    for f in ("group_nanmean", "group_nanvar", "group_nanstd"):
        _group_check(nb, f, v, labels, num_labels=K)


# ---------------------------------------------------------------------------- properties
def test_properties_large_fill(nb):
    import torch

    g = torch.Generator(device="cuda").manual_seed(0)
    x = torch.rand((1, 100_000_000), generator=g, device="cuda", dtype=torch.float64)
    x[x <= 0.3] = float("nan")
    x[0, :5] = float("nan")
    f = nb.ffill(x)
    # idempotence, no NaN after the first valid value, untouched valid values
    assert torch.equal(torch.nan_to_num(nb.ffill(f), nan=-1.0), torch.nan_to_num(f, nan=-1.0))
    first_valid = int(torch.nonzero(~torch.isnan(x[0]))[0])
    assert not torch.isnan(f[0, first_valid:]).any() and torch.isnan(f[0, :first_valid]).all()
    valid = ~torch.isnan(x)
    assert torch.equal(f[valid], x[valid])
    # bfill is ffill of the mirrored row
    b = nb.bfill(x)
    bm = nb.ffill(torch.flip(x, dims=[1]))
    assert torch.equal(torch.nan_to_num(torch.flip(bm, dims=[1]), nan=-1.0), torch.nan_to_num(b, nan=-1.0))
    # limit: every filled run is at most `limit` long
    fl = nb.ffill(x, limit=2)
    filled = (~torch.isnan(fl)) & torch.isnan(x)
    run3 = filled[0, 2:] & filled[0, 1:-1] & filled[0, :-2]
    assert not run3.any()


def test_properties_large_move_and_exp(nb):
    import torch

    g = torch.Generator(device="cuda").manual_seed(1)
    x = torch.rand((64, 1_000_000), generator=g, device="cuda", dtype=torch.float32)
    x[x <= 0.1] = float("nan")
    # window=1 mean is the identity (exact, reference test_moving.py:180-192)
    m1 = nb.move_mean(x, window=1, min_count=1)
    assert torch.equal(torch.nan_to_num(m1, nan=-1.0), torch.nan_to_num(x, nan=-1.0))
    # causality / tiling independence: results on a prefix equal the prefix of the results,
    # and shifting the row start (different tile alignment) changes nothing beyond rounding
    full = nb.move_std(x, window=1000, min_count=500)
    part = nb.move_std(x[:, :333_333].contiguous(), window=1000, min_count=500)
    torch.testing.assert_close(part, full[:, :333_333], rtol=1e-5, atol=1e-6, equal_nan=True)
    e_full = nb.move_exp_nanmean(x, alpha=np.float32(0.1))
    e_part = nb.move_exp_nanmean(x[:, :333_333].contiguous(), alpha=np.float32(0.1))
    torch.testing.assert_close(e_part, e_full[:, :333_333], rtol=1e-5, atol=1e-6, equal_nan=True)
    # mean * count == sum  (count from min_count-free move_sum of the validity mask)
    s = nb.move_sum(x, window=50, min_count=1)
    mean = nb.move_mean(x, window=50, min_count=1)
    cnt = nb.move_sum((~torch.isnan(x)).to(torch.float32), window=50, min_count=1)
    torch.testing.assert_close(mean * cnt, s, rtol=1e-5, atol=1e-5, equal_nan=True)


def test_properties_large_group(nb):
    import torch

    g = torch.Generator(device="cuda").manual_seed(2)
    rows, n, K = 2048, 200_000, 1000
    v = torch.rand((rows, n), generator=g, device="cuda", dtype=torch.float32)
    v[v <= 0.1] = float("nan")
    labels = torch.randint(-1, K, (n,), generator=g, device="cuda")
    s = nb.group_nansum(v, labels, num_labels=K, axis=-1)
    c = nb.group_nancount(v, labels, num_labels=K, axis=-1)
    keep = (labels >= 0)[None, :] & ~torch.isnan(v)
    # checksum of checksums: group sums add up to the masked row sums; counts are exact
    row_sum = torch.where(keep, v, torch.zeros_like(v)).to(torch.float64).sum(dim=1)
    torch.testing.assert_close(s.to(torch.float64).sum(dim=1), row_sum, rtol=1e-5, atol=0)
    assert torch.equal(c.sum(dim=1).to(torch.int64), keep.sum(dim=1))
    mean = nb.group_nanmean(v, labels, num_labels=K, axis=-1)
    torch.testing.assert_close(mean * c, s, rtol=1e-5, atol=1e-6, equal_nan=True)
    mx = nb.group_nanmax(v, labels, num_labels=K, axis=-1)
    am = nb.group_nanargmax(v, labels, num_labels=K, axis=-1)
    ok = ~torch.isnan(am)
    picked = torch.gather(v, 1, torch.nan_to_num(am, nan=0.0).to(torch.int64))
    assert torch.equal(picked[ok], mx[ok])


# ------------------------------------------ core-axis sharding protocol on a single device
def test_shard_protocol_halo_carry_partials_single_gpu(nb):
    """The C-ABI hooks that multi-GPU sharding uses (a_halo, carry_in/agg_out, group partial
    states + combine) driven from one device: two shards processed one after the other must
    reproduce the unsharded result."""
    import torch

    from numbagg_b200 import decorators as D
    from numbagg_b200 import distributed as nd

    a = fixture_array((4, 50_000), nan_frac=0.3, seed=31)
    b = a**2 + 1
    cut = 20_011  # odd cut: shard rows are not 16-byte aligned
    ta, tb = torch.from_numpy(a).cuda(), torch.from_numpy(b).cuda()
    s0a, s1a = ta[:, :cut].contiguous(), ta[:, cut:].contiguous()
    s0b, s1b = tb[:, :cut].contiguous(), tb[:, cut:].contiguous()
    for w in (20, 1000):
        for f in ("move_mean", "move_std", "move_corr"):
            arrs0, arrs1 = ([s0a, s0b], [s1a, s1b]) if f == "move_corr" else ([s0a], [s1a])
            halos = [x[:, -w:].contiguous() for x in arrs0]
            got = torch.cat([D.run_move(f, arrs0, w, 5, -1), D.run_move(f, arrs1, w, 5, -1, halos)], dim=1)
            args = (a, b) if f == "move_corr" else (a,)
            # corr divides by variances built from the reference's never-resynced running sums:
            # its own drift over 50k steps is ~1e-11 relative here (see test_move_vs_oracle_tiles)
            assert_parity(f, got.cpu().numpy(), getattr(oracle, f)(*args, window=w, min_count=5),
                          scale=_scale(f, list(args)), atol=1e-9 if f == "move_corr" else None)
    for f in EXP_ONE + EXP_TWO:
        arrs0, arrs1 = ([s0a, s0b], [s1a, s1b]) if f in EXP_TWO else ([s0a], [s1a])
        o0, agg0 = D.run_move_exp(f, arrs0, 0.05, 0.1, -1, None, True, True)
        o1, agg1 = D.run_move_exp(f, arrs1, 0.05, 0.1, -1, agg0, True, True)
        args = (a, b) if f in EXP_TWO else (a,)
        assert_parity(f, torch.cat([o0, o1], dim=1).cpu().numpy(), getattr(oracle, f)(*args, alpha=0.05, min_weight=0.1), scale=20.0)
        # aggregate-only pass + host-side composition == state after scanning both shards
        _, only1 = D.run_move_exp(f, arrs1, 0.05, 0.1, -1, None, True, False)
        torch.testing.assert_close(nd.exp_compose(f, agg0, only1), agg1, rtol=1e-12, atol=1e-300)
    x = a.copy()
    x[1, 15_000:30_000] = np.nan
    tx = torch.from_numpy(x).cuda()
    for limit in (3, 20_000, 50_000):
        f0, g0 = D.run_fill("ffill", tx[:, :cut].contiguous(), limit, -1, None, True, True)
        f1, _ = D.run_fill("ffill", tx[:, cut:].contiguous(), limit, -1, g0, False, True)
        np.testing.assert_array_equal(torch.cat([f0, f1], dim=1).cpu().numpy(), oracle.ffill(x, limit=limit))
        b1, h1 = D.run_fill("bfill", tx[:, cut:].contiguous(), limit, -1, None, True, True)
        b0, _ = D.run_fill("bfill", tx[:, :cut].contiguous(), limit, -1, h1, False, True)
        np.testing.assert_array_equal(torch.cat([b0, b1], dim=1).cpu().numpy(), oracle.bfill(x, limit=limit))
    labels = np.random.RandomState(32).randint(-1, 40, size=a.shape[1])
    tl = torch.from_numpy(labels).cuda()
    for f in oracle.GROUPED_FUNCS:
        p0 = D.run_group_partial(f, s0a, tl[:cut].contiguous(), 40, 0)
        p1 = D.run_group_partial(f, s1a, tl[cut:].contiguous(), 40, cut)
        how = nd._GROUP_COMBINE.get(f, {})
        if how and all(h in ("sum_v", "sum_i") for h in how.values()):
            # additive ops: what the all-reduce does, on the channel views of the two states
            c0 = D.group_state_channels(f, p0, s0a.shape[0], 40)
            c1 = D.group_state_channels(f, p1, s0a.shape[0], 40)
            for ch, h in how.items():
                if h == "sum_v":
                    c0[ch].copy_((c0[ch].contiguous().view(torch.float64) + c1[ch].contiguous().view(torch.float64)).view(torch.int64))
                else:
                    c0[ch].copy_(c0[ch] + c1[ch])
        else:
            D.run_group_combine(f, np.float64, p0, p1)
        got = D.run_group_finalize(f, np.float64, p0, 1).cpu().numpy()
        _group_check_arrays(f, got, getattr(oracle, f)(a, labels, num_labels=40, axis=-1), a)


def _group_check_arrays(f, got, exp, values):
    m = float(np.nanmax(np.abs(values)))
    scale = m * m * 1500 if (f in GROUP_FLOAT_ONLY or f == "group_nansum_of_squares") else (m * 1500 if f in ("group_nansum", "group_nanmean") else None)
    assert_parity(f, got, exp, scale=scale)


def test_numpy_inputs_pipelined_in_row_blocks(nb):
    """Large numpy inputs are streamed through the GPU in row blocks (H2D / kernels / D2H
    overlapped on several streams): same values as the one-shot path and the oracle."""
    a = fixture_array((96, 200_000), nan_frac=0.2, seed=41)  # 154 MB > the pipelining threshold
    b = a**2 + 1
    pa, pb = nb.empty_pinned(a.shape, a.dtype), nb.empty_pinned(a.shape, a.dtype)
    pa[...], pb[...] = a, b
    for arr_a, arr_b in ((a, b), (pa, pb)):
        assert_parity("move_mean", nb.move_mean(arr_a, window=20, min_count=1), oracle.move_mean(a, window=20, min_count=1), scale=1.0)
        assert_parity("move_corr", nb.move_corr(arr_a, arr_b, window=50, min_count=5), oracle.move_corr(a, b, window=50, min_count=5), scale=1.0, atol=1e-9)
        assert_parity("move_exp_nanmean", nb.move_exp_nanmean(arr_a, alpha=0.1), oracle.move_exp_nanmean(a, alpha=0.1), scale=1.0)
        np.testing.assert_array_equal(nb.ffill(arr_a, limit=3), oracle.ffill(a, limit=3))
        np.testing.assert_array_equal(nb.bfill(arr_a), oracle.bfill(a))
    a3 = a.reshape(96, 400, 500)
    assert_parity("move_sum", nb.move_sum(a3, window=7, min_count=1, axis=1), oracle.move_sum(a3, window=7, min_count=1, axis=1), scale=7.0)


def test_time_major_layout_few_columns(nb):
    """(time, few series) C-ordered arrays with axis=0: the core axis is strided and there
    are too few columns for one-thread-per-column kernels, so it is transposed once and the
    row-tile kernels run along it.  Values must not depend on the route taken."""
    a = fixture_array((40_000, 6), nan_frac=0.3, seed=51)
    b = a**2 + 1
    assert_parity("move_mean", nb.move_mean(a, window=50, min_count=5, axis=0), oracle.move_mean(a, window=50, min_count=5, axis=0), scale=1.0)
    assert_parity("move_cov", nb.move_cov(a, b, window=50, min_count=5, axis=0), oracle.move_cov(a, b, window=50, min_count=5, axis=0), scale=4.0)
    assert_parity("move_exp_nanvar", nb.move_exp_nanvar(a, alpha=0.05, axis=0), oracle.move_exp_nanvar(a, alpha=0.05, axis=0), scale=1.0)
    al = np.random.RandomState(52).rand(40_000, 6) * 0.5 + 0.01
    assert_parity("move_exp_nanmean", nb.move_exp_nanmean(a, alpha=al, axis=0), oracle.move_exp_nanmean(a, alpha=al, axis=0), scale=1.0)
    assert_parity("move_exp_nansum", nb.move_exp_nansum(a, alpha=al[:, 0].copy(), axis=0), oracle.move_exp_nansum(a, alpha=al[:, 0].copy(), axis=0), scale=100.0)
    for limit in (None, 4):
        np.testing.assert_array_equal(nb.ffill(a, limit=limit, axis=0), oracle.ffill(a, limit=limit, axis=0))
        np.testing.assert_array_equal(nb.bfill(a, limit=limit, axis=0), oracle.bfill(a, limit=limit, axis=0))
    a3 = fixture_array((3, 5000, 4), seed=53)
    assert_parity("move_sum", nb.move_sum(a3, window=9, min_count=1, axis=1), oracle.move_sum(a3, window=9, min_count=1, axis=1), scale=9.0)


def test_numpy_1d_inputs_pipelined_along_core_axis(nb):
    """Long 1-D numpy inputs are streamed in core-axis chunks with the state (halo / carry)
    handed from chunk to chunk on the device."""
    n = 30_000_000  # 240 MB float64 -> 2 chunks
    a = fixture_array((n,), nan_frac=0.3, seed=61)
    a[14_998_000:15_002_000] = np.nan  # a NaN run across the chunk boundary (shorter than the
    # ~7070 steps after which 0.9**k underflows: beyond that the reference returns ratios of
    # stuck subnormals -- DESIGN.md "known parity limits")
    assert_parity("move_exp_nanmean", nb.move_exp_nanmean(a, alpha=0.1), oracle.move_exp_nanmean(a, alpha=0.1), scale=1.0)
    # (slow decays are kept to ~1e4-term memories: with alpha=1e-7 two valid evaluation orders of
    # the 1e7-term recurrence already differ by 2e-11 relative)
    assert_parity("move_exp_nansum", nb.move_exp_nansum(a, alpha=1e-4, min_weight=1e-4), oracle.move_exp_nansum(a, alpha=1e-4, min_weight=1e-4), scale=1e4)
    assert_parity("move_exp_nanmean", nb.move_exp_nanmean(a.reshape(1, 1, n), alpha=0.1), oracle.move_exp_nanmean(a.reshape(1, 1, n), alpha=0.1), scale=1.0)
    a[9_000_000:16_000_000] = np.nan  # fills: a NaN run of 7M elements across the chunk boundary
    for limit in (None, 5, 6_500_000):
        np.testing.assert_array_equal(nb.ffill(a, limit=limit), oracle.ffill(a, limit=limit))
        np.testing.assert_array_equal(nb.bfill(a, limit=limit), oracle.bfill(a, limit=limit))
    b = a**2 + 1
    assert_parity("move_mean", nb.move_mean(a, window=1000, min_count=10), oracle.move_mean(a, window=1000, min_count=10), scale=1.0)
    assert_parity("move_cov", nb.move_cov(a, b, window=50, min_count=5), oracle.move_cov(a, b, window=50, min_count=5), scale=4.0)
