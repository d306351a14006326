"""Generate tests/golden/*.npz from the reference's own Numba path.

Run in the dev container only (needs /root/reference and numba):

    PYTHONDONTWRITEBYTECODE=1 python oracle/gen_golden.py

The reference cannot travel to the GPU box, so its outputs are frozen here as small
fixtures.  Each case stores the input arrays, the keyword arguments (JSON) and the output of
``numbagg.<func>(*args, **kwargs)``.  tests/test_oracle_golden.py replays every case through
the C oracle (bit-exact) and tests/test_gpu_golden.py through the CUDA path (north_star
tolerances).  This script is test infrastructure, not product.
"""

from __future__ import annotations

import hashlib
import json
import os
import sys

import numpy as np

REF = os.environ.get("NUMBAGG_REFERENCE", "/root/reference")
sys.path.insert(0, REF)
import numbagg  # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))
OUT = os.path.join(os.path.dirname(HERE), "tests", "golden")


def fixture_array(shape, nan_frac=0.1, seed=0, dtype=np.float64):
    """The reference's shared test fixture (numbagg/test/conftest.py:581-592)."""
    a = np.random.RandomState(seed).rand(*shape)
    a = np.where(a > nan_frac, a, np.nan)
    return a.astype(dtype)


class Suite:
    def __init__(self, name):
        self.name = name
        self.arrays = {}
        self.manifest = []
        self._by_hash = {}

    def _store(self, arr) -> str:
        """Store each distinct array once (inputs are shared by many cases)."""
        arr = np.asarray(arr)
        h = hashlib.sha1(arr.tobytes() + str((arr.dtype, arr.shape)).encode()).hexdigest()
        if h not in self._by_hash:
            key = f"x{len(self._by_hash)}"
            self._by_hash[h] = key
            self.arrays[key] = arr
        return self._by_hash[h]

    def add(self, func, args, kwargs=None, note="", layout=None):
        """layout: memory order of args[0] as a permutation of its axes (npz files store
        C-order only; tests/_golden.py re-creates the strides before replaying the case)."""
        kwargs = kwargs or {}
        if layout is not None:
            args = [np.ascontiguousarray(args[0].transpose(layout)).transpose(np.argsort(layout))] + list(args[1:])
        idx = len(self.manifest)
        arr_kwargs = {}
        plain_kwargs = {}
        for k, v in kwargs.items():
            if isinstance(v, np.ndarray):
                arr_kwargs[k] = self._store(v)
            elif isinstance(v, np.generic):
                plain_kwargs[k] = {"np_scalar": str(v.dtype), "value": float(v)}
            elif isinstance(v, tuple):
                plain_kwargs[k] = {"tuple": list(v)}
            else:
                plain_kwargs[k] = v
        with np.errstate(all="ignore"):
            out = getattr(numbagg, func)(*args, **kwargs)
        arg_keys = []
        for a in args:
            arg_keys.append(self._store(a))
        self.manifest.append(
            dict(func=func, args=arg_keys, kwargs=plain_kwargs, array_kwargs=arr_kwargs,
                 out=self._store(out), note=note, **({"layout": list(layout)} if layout is not None else {}))
        )

    def save(self):
        os.makedirs(OUT, exist_ok=True)
        self.arrays["manifest"] = np.frombuffer(json.dumps(self.manifest).encode(), dtype=np.uint8)
        path = os.path.join(OUT, f"{self.name}.npz")
        np.savez_compressed(path, **self.arrays)
        print(f"{path}: {len(self.manifest)} cases, {os.path.getsize(path) / 1024:.0f} KiB")


nan = np.nan


def gen_moving():
    s = Suite("moving")
    one = ["move_mean", "move_sum", "move_std", "move_var"]
    two = ["move_cov", "move_corr"]
    # reference test grid: test_moving.py:29-53 (window x min_count on (3,500)), both dtypes
    for dtype in (np.float64, np.float32):
        a = fixture_array((3, 200), dtype=dtype)
        b = (a**2 + 1).astype(dtype)  # conftest.py:73-77
        for window in (1, 10, 50):
            for min_count in (None, 0, 1, 3, window):
                for f in one:
                    s.add(f, [a], dict(window=window, min_count=min_count))
                for f in two:
                    s.add(f, [a, b], dict(window=window, min_count=min_count))
    # other axes / ndim
    a3 = fixture_array((6, 40, 5), seed=1)
    for axis in (0, 1, 2, -2):
        for f in one:
            s.add(f, [a3], dict(window=4, min_count=2, axis=axis))
        for f in two:
            s.add(f, [a3, a3**2 + 1], dict(window=4, min_count=2, axis=axis))
    # window == n, long single row crossing many GPU tiles, heavy NaN
    long = fixture_array((9000,), nan_frac=0.3, seed=2)
    for f in one:
        s.add(f, [long], dict(window=9000, min_count=1))
        s.add(f, [long], dict(window=1000, min_count=500))
        s.add(f, [long], dict(window=20, min_count=1))
    for f in two:
        s.add(f, [long, long**2 + 1], dict(window=1000, min_count=500))
    # SURVEY appendix A vectors
    b7 = np.array([1, nan, 3, 4, nan, nan, 7.0])
    for mc in (0, 1, None):
        for f in one:
            s.add(f, [b7], dict(window=3, min_count=mc))
        for f in two:
            s.add(f, [b7, 2 * b7], dict(window=3, min_count=mc))
    sparse = np.array([nan, nan, nan, 1, nan, nan, nan, nan])
    for f in one:
        s.add(f, [sparse], dict(window=2, min_count=0))
    s.add("move_sum", [np.ones(5)], dict(window=2, min_count=3), "min_count > window is legal")
    s.add("move_mean", [np.arange(4.0)], dict(window=4, min_count=1))
    s.add("move_mean", [np.arange(10)], dict(window=3), "int input -> float64")
    s.add("move_mean", [np.arange(10, dtype=np.float16)], dict(window=3), "float16 -> float32")
    # float32 stability tests test_moving.py:180-192
    arr = np.array([0.1, 0.2, 0.3] * 100, dtype=np.float32)
    s.add("move_mean", [arr], dict(window=1))
    s.add("move_sum", [arr], dict(window=1))
    s.add("move_sum", [np.tile(np.arange(10, dtype=np.float32) * 1.7, 30)], dict(window=10))
    # inf handling
    s.add("move_sum", [np.array([1.0, np.inf, 2.0, -np.inf, 3.0, 4.0, 5.0])], dict(window=2, min_count=1))
    s.save()


def gen_moving_exp():
    s = Suite("moving_exp")
    one = ["move_exp_nancount", "move_exp_nanmean", "move_exp_nansum", "move_exp_nanvar", "move_exp_nanstd"]
    two = ["move_exp_nancov", "move_exp_nancorr"]
    for dtype in (np.float64, np.float32):
        a = fixture_array((3, 200), dtype=dtype)
        b = (a**2 + 1).astype(dtype)
        for alpha in (0.5, 0.1):
            al = dtype(alpha) if dtype is np.float32 else alpha
            for mw in (0, 0.3):
                for f in one:
                    s.add(f, [a], dict(alpha=al, min_weight=mw))
                for f in two:
                    s.add(f, [a, b], dict(alpha=al, min_weight=mw))
    # float32 data + python float alpha -> float64 loop
    a32 = fixture_array((2, 100), dtype=np.float32)
    s.add("move_exp_nanmean", [a32], dict(alpha=0.25), "f32 + py float alpha -> f64")
    # alpha as arrays
    a = fixture_array((4, 150), seed=3)
    al1 = np.random.RandomState(5).rand(150) * 0.9 + 0.05
    aln = np.random.RandomState(6).rand(4, 150) * 0.9 + 0.05
    for f in one:
        s.add(f, [a], dict(alpha=al1))
        s.add(f, [a], dict(alpha=aln))
        s.add(f, [a.T.copy()], dict(alpha=aln.T.copy(), axis=0))
        s.add(f, [a.T.copy()], dict(alpha=al1, axis=0))
    for f in two:
        s.add(f, [a, a**2 + 1], dict(alpha=al1))
        s.add(f, [a, a**2 + 1], dict(alpha=aln))
    # long rows crossing many tiles; 30 % NaN (config 3 style)
    long = fixture_array((9000,), nan_frac=0.3, seed=7)
    for f in one:
        s.add(f, [long], dict(alpha=0.1))
        s.add(f, [long], dict(alpha=0.001, min_weight=0.5))
    for f in two:
        s.add(f, [long, long**2 + 1], dict(alpha=0.1))
    # known-answer inputs of test_moving_exp.py
    d = np.array([nan, 1, nan, 2, 3.0])
    for f in one:
        s.add(f, [d], dict(alpha=0.5))
    s.add("move_exp_nanmean", [d], dict(alpha=0.5, min_weight=0.6))
    for f in two:
        s.add(f, [d, d * d], dict(alpha=0.5))
    s.add("move_exp_nancount", [np.array([1, 0, nan, nan, 1, 0.0])], dict(alpha=0.5))
    s.add("move_exp_nancount", [np.array([1, 0, nan, nan, 1, 0.0])], dict(alpha=0.25))
    s.add("move_exp_nanmean", [np.array([10, 0, nan, 10.0])], dict(alpha=0.5))
    s.add("move_exp_nanmean", [np.array([10, 0, nan, 10.0])], dict(alpha=0.25))
    s.add("move_exp_nansum", [np.array([10, 0, nan, 10.0])], dict(alpha=0.5))
    s.add("move_exp_nansum", [np.array([10, 0, nan, 10.0])], dict(alpha=0.25))
    s.add("move_exp_nancorr", [np.array([10, 0, 5, 10.0]), np.array([10, 0, 10, 5.0])], dict(alpha=0.5))
    s.add("move_exp_nancorr", [np.array([10, 0, 5, 10.0]), np.array([10, 0, 10, 5.0])], dict(alpha=0.25))
    # min_weight counts (test_moving_exp.py:44-79)
    arr = np.ones(25)
    arr[:5] = nan
    for f in one:
        for mw in (0.0, 0.5, 0.9, 1.0):
            for alpha in (0.2, 0.8):
                s.add(f, [arr], dict(alpha=alpha, min_weight=mw))
    # NaN-mask patterns (test_moving_exp.py:160-236)
    pats = [
        [nan, nan], [5.0, nan], [1.0, nan, 2.0], [1.0, nan, 1.0], [1.0, nan], [0.1, nan],
        [0.75, nan], [0.5, nan], [0.9, nan], [0.95, nan, 1.0],
        [0.59288027, nan, 0.4758262, 0.70877039],
    ]
    for p in pats:
        p = np.array(p)
        for alpha in (0.1, 0.5, 0.9):
            for f in ("move_exp_nanvar", "move_exp_nanstd"):
                s.add(f, [p], dict(alpha=alpha))
            for f in two:
                s.add(f, [p, p], dict(alpha=alpha))
    # inf / alpha=1 / big values (test_moving_exp.py:251-295)
    s.add("move_exp_nanmean", [np.array([np.inf])], dict(alpha=0.25))
    s.add("move_exp_nanmean", [np.array([0, 0, np.inf], dtype=np.float16)], dict(alpha=1.0))
    s.add("move_exp_nanmean", [np.array([0, np.inf, np.inf], dtype=np.float16)], dict(alpha=1.0))
    big = np.array([[0.0, 1.19846209e308], [1.19846209e308, 1.19846209e308]])
    s.add("move_exp_nanmean", [big], dict(alpha=1.0))
    s.add("move_exp_nanmean", [np.array([[0, nan]])], dict(alpha=1.0))
    s.add("move_exp_nanmean", [np.array([[1, nan]])], dict(alpha=1.0))
    # all-NaN count (test_moving_exp.py:133-157)
    for alpha in (0.1, 0.5, 0.9):
        s.add("move_exp_nancount", [np.array([nan, nan, 1.0, nan, 1.0])], dict(alpha=alpha, min_weight=0.0))
        s.add("move_exp_nancount", [np.array([nan] * 4)], dict(alpha=alpha, min_weight=0.0))
        s.add("move_exp_nansum", [np.array([nan] * 4)], dict(alpha=alpha, min_weight=0.0))
    s.save()


def gen_fill():
    s = Suite("fill")
    a10 = np.array([nan, 1, nan, nan, nan, 2, nan, 3, nan, nan])
    for f in ("ffill", "bfill"):
        for limit in (None, 0, 1, 2, 3):
            s.add(f, [a10], dict(limit=limit))
        for dtype in (np.float64, np.float32):
            a = fixture_array((3, 200), nan_frac=0.4, dtype=dtype)
            for limit in (None, 1, 3):
                s.add(f, [a], dict(limit=limit))
                s.add(f, [a], dict(limit=limit, axis=0))
        long = fixture_array((12000,), nan_frac=0.3, seed=7)
        long[3000:9000] = nan  # a NaN run longer than any GPU tile
        for limit in (None, 5, 4000):
            s.add(f, [long], dict(limit=limit))
        s.add(f, [np.array([np.inf, nan, -np.inf, nan])], {})
        s.add(f, [np.arange(5, dtype=np.int32)], {})
        s.add(f, [np.arange(5, dtype=np.int64)], dict(limit=1))
        s.add(f, [np.full(7, nan)], {})
    s.save()


def gen_grouped():
    s = Suite("grouped")
    funcs = [
        "group_nanmean", "group_nansum", "group_nancount", "group_nanargmax", "group_nanargmin",
        "group_nanfirst", "group_nanlast", "group_nanprod", "group_nansum_of_squares",
        "group_nanvar", "group_nanstd", "group_nanmin", "group_nanmax", "group_nanany",
        "group_nanall",
    ]
    float_only = {"group_nanvar", "group_nanstd"}
    rs = np.random.RandomState(0)
    # appendix A vector
    v = np.array([1, nan, 3, 2, 3.0])
    lab = np.array([0, 0, 2, 2, 2])
    for f in funcs:
        s.add(f, [v, lab], dict(num_labels=4))
    s.add("group_nanvar", [np.array([1.0, 2, 4]), np.zeros(3, dtype=np.int64)], dict(ddof=0))
    s.add("group_nanvar", [np.array([1.0, 2, 4]), np.zeros(3, dtype=np.int64)], dict(ddof=1))
    s.add("group_nanargmax", [np.array([1, 3, 3, nan]), np.zeros(4, dtype=np.int64)], {}, "ties -> first")
    s.add("group_nanargmin", [np.array([3, 1, 1, nan]), np.zeros(4, dtype=np.int64)], {}, "ties -> first")
    # reference fixture style: 12 random labels (conftest.py:164-172), -1 = missing, an
    # empty group (test_grouped.py:87-108)
    for dtype in (np.float64, np.float32):
        vals = fixture_array((2000,), dtype=dtype, seed=11)
        labels = rs.randint(-1, 12, size=2000)
        labels[labels == 7] = 3  # group 7 is empty
        for f in funcs:
            s.add(f, [vals, labels], dict(num_labels=13))
        # zeros / negative values so any/all/prod/min/max see sign structure
        vals2 = np.round((fixture_array((600,), dtype=dtype, seed=12) - 0.5) * 6)
        lab2 = rs.randint(0, 5, size=600)
        for f in funcs:
            s.add(f, [vals2, lab2], {})
        # 2-D, axis forms
        v2 = fixture_array((7, 300), dtype=dtype, seed=13)
        l_last = rs.randint(0, 6, size=300)
        l_first = rs.randint(0, 3, size=7)
        l_full = rs.randint(-1, 9, size=(7, 300))
        for f in funcs:
            s.add(f, [v2, l_last], dict(axis=-1))
            s.add(f, [v2, l_first], dict(axis=0))
            s.add(f, [v2, l_full], dict(axis=None))
        v3 = fixture_array((4, 6, 50), dtype=dtype, seed=14)
        l23 = rs.randint(0, 5, size=(6, 50))
        for f in funcs:
            s.add(f, [v3, l23], dict(axis=(1, 2)))
    # integer / bool values (test_grouped.py:359-394, 593-612); narrow label dtypes (434-473)
    for dtype in (np.int32, np.int64):
        iv = rs.randint(-5, 6, size=400).astype(dtype)
        il = rs.randint(0, 6, size=400)
        for f in funcs:
            if f in float_only:
                continue
            s.add(f, [iv, il], dict(num_labels=6))
    bv = rs.rand(200) > 0.5
    bl = rs.randint(0, 4, size=200)
    for f in ("group_nansum", "group_nanany", "group_nanall", "group_nancount", "group_nanmean"):
        s.add(f, [bv, bl], {})
    for ldt in (np.int8, np.int16, np.int32):
        s.add("group_nansum", [fixture_array((300,), seed=15), rs.randint(0, 5, size=300).astype(ldt)], {})
    # ddof variants
    vals = fixture_array((500,), seed=16)
    labels = rs.randint(0, 8, size=500)
    for ddof in (0, 1, 2):
        s.add("group_nanvar", [vals, labels], dict(ddof=ddof))
        s.add("group_nanstd", [vals, labels], dict(ddof=ddof))
    # many labels (beyond a shared-memory bin table) with collisions
    vals = fixture_array((30000,), seed=17)
    labels = rs.randint(0, 20000, size=30000)
    for f in funcs:
        s.add(f, [vals, labels], dict(num_labels=20000))
    s.save()


def gen_reduce():
    """Plain NaN reductions (numbagg/funcs.py:23-242): SURVEY 8(f) rank 1."""
    s = Suite("reduce")
    funcs = ["allnan", "anynan", "nancount", "nansum", "nanmean", "nanvar", "nanstd",
             "nanargmax", "nanargmin", "nanmax", "nanmin"]
    float_only = {"nanmean", "nanvar", "nanstd"}
    raises_on_allnan = {"nanargmax", "nanargmin"}
    rs = np.random.RandomState(5)
    _add = s.add

    def add(f, args, kwargs=None, note="", layout=None):
        try:
            _add(f, args, kwargs, note, layout)
        except ValueError as e:  # nanarg* on a slice without data: covered by the error tests
            assert "All-NaN" in str(e), e

    s.add = add
    for dtype in (np.float64, np.float32):
        a1 = fixture_array((2000,), dtype=dtype, seed=21)
        a2 = fixture_array((7, 300), nan_frac=0.3, dtype=dtype, seed=22)
        a3 = fixture_array((4, 6, 50), dtype=dtype, seed=23)
        signed = np.round((fixture_array((5, 400), dtype=dtype, seed=24) - 0.5) * 8)  # ties, zeros, negatives
        withrow = a2.copy()
        withrow[2] = nan  # an all-NaN row
        special = np.array([[np.inf, nan, 1.0, -np.inf], [nan, -np.inf, nan, -np.inf],
                            [0.0, -0.0, nan, 0.0], [3.0, 3.0, 3.0, 3.0]], dtype=dtype)
        big = (fixture_array((3, 30000), dtype=dtype, seed=25) * 1e3 + 1e6).astype(dtype)  # mean >> spread
        for f in funcs:
            s.add(f, [a1], {})
            s.add(f, [a1], dict(axis=0))
            for axis in (None, -1, 0):
                s.add(f, [a2], dict(axis=axis))
                s.add(f, [signed], dict(axis=axis))
            for axis in (None, 0, 1, 2, (0, 1), (1, 2), (0, 2), (2, 0)):
                s.add(f, [a3], dict(axis=axis))
            for axis in (None, 1, (0, 1), (1, 2)):
                s.add(f, [a3], dict(axis=axis), "permuted memory layout", layout=(2, 0, 1))
            s.add(f, [big], dict(axis=-1))
            if f not in raises_on_allnan:
                s.add(f, [withrow], dict(axis=-1), "all-NaN row")
                s.add(f, [special], dict(axis=-1), "inf / signed zero / constant rows")
            else:
                s.add(f, [special[[0, 2, 3]]], dict(axis=-1), "inf / signed zero / constant rows")
                s.add(f, [special[[0, 1, 3]]], dict(axis=0))
        for f in ("nanvar", "nanstd"):
            for ddof in (0, 1, 2, 5):
                s.add(f, [a2], dict(axis=-1, ddof=ddof))
            s.add(f, [a2[:, :2]], dict(axis=-1, ddof=2), "count <= ddof")
    for dtype in (np.int32, np.int64):
        i2 = rs.randint(-50, 50, size=(6, 250)).astype(dtype)
        i3 = rs.randint(-5, 6, size=(3, 5, 40)).astype(dtype)
        wide = rs.randint(np.iinfo(dtype).min // 4, np.iinfo(dtype).max // 4, size=(4, 300)).astype(dtype)
        for f in funcs:
            if f in float_only:
                continue
            for axis in (None, -1, 0):
                s.add(f, [i2], dict(axis=axis))
            for axis in (1, (0, 2), (2, 1)):
                s.add(f, [i3], dict(axis=axis))
            s.add(f, [wide], dict(axis=-1), "wide integers (int64 compares as float64)")
    # integers through the float-only loops (NumPy casts to float64)
    ii = rs.randint(-9, 10, size=(5, 60)).astype(np.int64)
    for f in sorted(float_only):
        s.add(f, [ii], dict(axis=-1))
    s.save()


def gen_quantile():
    """nanquantile / nanmedian (numbagg/funcs.py:245-291, 332-335): SURVEY 8(f) rank 3."""
    s = Suite("quantile")
    rs = np.random.RandomState(9)
    qs = [0.5, 0.0, 1.0, [0.25, 0.5, 0.75], [0.1, nan, 0.9], [0.999, 0.001, 0.5, 0.5]]
    for dtype in (np.float64, np.float32):
        a1 = fixture_array((3000,), dtype=dtype, seed=31)
        a2 = fixture_array((7, 300), nan_frac=0.4, dtype=dtype, seed=32)
        a3 = fixture_array((4, 6, 50), dtype=dtype, seed=33)
        ties = np.round((fixture_array((5, 400), dtype=dtype, seed=34) - 0.5) * 6)
        long = rs.standard_normal((2, 20000)).astype(dtype)  # longer than one shared-memory sort
        long[0, ::7] = nan
        for q in qs:
            s.add("nanquantile", [a1], dict(quantiles=q))
            for axis in (-1, 0, None):
                s.add("nanquantile", [a2], dict(quantiles=q, axis=axis))
                s.add("nanquantile", [ties], dict(quantiles=q, axis=axis))
            for axis in (1, (0, 2), (2, 1), None):
                s.add("nanquantile", [a3], dict(quantiles=q, axis=axis))
            s.add("nanquantile", [long], dict(quantiles=q, axis=-1))
        for axis in (-1, 0, None):
            s.add("nanmedian", [a2], dict(axis=axis))
    special = np.array([[np.inf, 1, 2, nan], [np.inf, np.inf, -np.inf, 0], [nan] * 4, [3, 3, 3, 3.0], [-0.0, 0.0, -1, 1]])
    for q in ([0, 0.5, 1], 0.3):
        s.add("nanquantile", [special], dict(quantiles=q, axis=-1), "inf / all-NaN / constant rows")
    ii = rs.randint(-50, 50, size=(6, 250))
    s.add("nanquantile", [ii], dict(quantiles=[0.2, 0.8], axis=-1), "integers are cast to float64")
    s.add("nanquantile", [np.arange(10.0)], dict(quantiles=0.45))
    s.add("nanquantile", [np.array([5.0])], dict(quantiles=[0, 0.5, 1]))
    s.save()


def gen_matrix():
    """Matrix functions (funcs.py:338-532, moving_matrix.py:16-432): SURVEY 8(f) rank 2.
    Oracle-only so far: these fixtures pin oracle/nbg_oracle.c ahead of the CUDA kernels."""
    s = Suite("matrix")
    rs = np.random.RandomState(13)
    for dtype in (np.float64, np.float32):
        for nan_frac in (0.0, 0.25, 0.7):
            vo = fixture_array((5, 80), nan_frac=nan_frac, dtype=dtype, seed=41)      # (vars, obs)
            vo[3] = vo[1] * 2 + 1            # a perfectly correlated pair
            vo[4, :] = dtype(0.75)           # a constant variable: zero variance
            b3 = fixture_array((2, 3, 4, 50), nan_frac=nan_frac, dtype=dtype, seed=42)  # batch dims
            for f in ("nancorrmatrix", "nancovmatrix"):
                s.add(f, [vo])
                s.add(f, [b3])
            ov = np.ascontiguousarray(vo.T)   # (obs, vars)
            ob = np.ascontiguousarray(np.swapaxes(b3, -1, -2))
            for f in ("move_corrmatrix", "move_covmatrix"):
                for window, mc in ((10, None), (10, 3), (1, 1), (2, 1), (80, 5), (25, 1)):
                    s.add(f, [ov], dict(window=window, min_count=mc))
                s.add(f, [ob], dict(window=7, min_count=2))
            al = (rs.rand(80) * 0.8 + 0.1).astype(dtype)
            for f in ("move_exp_nancorrmatrix", "move_exp_nancovmatrix"):
                for alpha in (0.1, 0.9, np.float32(0.25), al):
                    for mw in (0, 0.3):
                        s.add(f, [ov], dict(alpha=alpha, min_weight=mw))
                s.add(f, [ob], dict(alpha=0.2))
        one = fixture_array((1, 30), dtype=dtype, seed=43)
        s.add("nancorrmatrix", [one])
        s.add("nancovmatrix", [one])
        s.add("move_covmatrix", [np.ascontiguousarray(one.T)], dict(window=5, min_count=1))
    ii = rs.randint(-5, 6, size=(4, 40))
    s.add("nancovmatrix", [ii], {}, "integers go through the float64 loop")
    s.add("move_corrmatrix", [np.ascontiguousarray(ii.T)], dict(window=6, min_count=2))
    s.save()


if __name__ == "__main__":
    gen_matrix()
    gen_quantile()
    gen_reduce()
    gen_moving()
    gen_moving_exp()
    gen_fill()
    gen_grouped()
