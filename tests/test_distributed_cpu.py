"""world_size-2 gloo tests of numbagg_b200.distributed on CPU: the exchange logic (halo
assembly, carry folding in rank order, ordered merge / all-reduce of grouped partial states)
must reproduce the unsharded result.  Local compute is the oracle-backed stand-in of
tests/_cpu_backend.py; on the GPU box the same code paths run with the CUDA kernels
(tests/test_distributed_gpu.py)."""

import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from oracle import oracle


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def fixture_array(shape, nan_frac=0.2, seed=0):
    a = np.random.RandomState(seed).rand(*shape)
    return np.where(a > nan_frac, a, np.nan)


def _split(n, world, uneven):
    cuts = [0]
    for r in range(world):
        cuts.append(n * (r + 1) // world if not uneven else min(n, cuts[-1] + (3 if r == 0 else n)))
    cuts[-1] = n
    return cuts


def _worker(rank, world, port, case, uneven, results):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from numbagg_b200 import distributed as nd
        from tests._cpu_backend import OracleBackend as B

        a = fixture_array((3, 60), seed=1)
        b = a**2 + 1
        cuts = _split(60, world, uneven)
        lo, hi = cuts[rank], cuts[rank + 1]
        lens = [cuts[r + 1] - cuts[r] for r in range(world)]
        ta, tb = torch.from_numpy(a[:, lo:hi].copy()), torch.from_numpy(b[:, lo:hi].copy())
        out = {}
        if case == "move":
            for f in ("move_mean", "move_sum", "move_std", "move_var"):
                out[f] = nd.move_sharded(f, ta, window=7, min_count=2, axis=-1, shard_lens=lens, backend=B).numpy()
            for f in ("move_cov", "move_corr"):
                out[f] = nd.move_sharded(f, ta, tb, window=7, min_count=2, axis=-1, shard_lens=lens, backend=B).numpy()
            # core axis 0: shards are row blocks of the transposed problem
            tt = torch.from_numpy(a.T[lo:hi].copy())
            out["move_mean_axis0"] = nd.move_sharded("move_mean", tt, window=7, min_count=2, axis=0, shard_lens=lens, backend=B).numpy().T
        elif case == "exp":
            for f in ("move_exp_nancount", "move_exp_nanmean", "move_exp_nansum", "move_exp_nanvar", "move_exp_nanstd"):
                out[f] = nd.move_exp_sharded(f, ta, alpha=0.2, min_weight=0.1, axis=-1, backend=B).numpy()
            for f in ("move_exp_nancov", "move_exp_nancorr"):
                out[f] = nd.move_exp_sharded(f, ta, tb, alpha=0.2, axis=-1, backend=B).numpy()
            al = np.random.RandomState(3).rand(60) * 0.8 + 0.1
            out["move_exp_nanmean_alpha1d"] = nd.move_exp_sharded(
                "move_exp_nanmean", ta, alpha=torch.from_numpy(al[lo:hi].copy()), axis=-1, backend=B).numpy()
        elif case == "exp_single":
            # long enough for the ONE-pass form: zero-carry scan + aggregate, then only the head
            # (exp_forget_length) is recomputed with the carry
            n2 = 2400
            x = fixture_array((2, n2), seed=11)
            x[1, 1100:1300] = np.nan  # a NaN run across the boundary
            c2 = _split(n2, world, uneven)
            px = torch.from_numpy(x[:, c2[rank]:c2[rank + 1]].copy())
            assert nd.exp_forget_length(0.9999) * 4 <= px.shape[1] or uneven
            for f in ("move_exp_nanmean", "move_exp_nansum", "move_exp_nanvar"):
                out[f] = nd.move_exp_sharded(f, px, alpha=0.9999, axis=-1, backend=B).numpy()
        elif case == "fill":
            x = a.copy()
            x[1, 10:50] = np.nan  # a NaN run across the shard boundary
            x[2, :] = np.nan
            tx = torch.from_numpy(x[:, lo:hi].copy())
            for f in ("ffill", "bfill"):
                for limit in (None, 2, 25):
                    out[f"{f}_{limit}"] = nd.fill_sharded(f, tx, limit=limit, axis=-1, shard_lens=lens, backend=B).numpy()
        elif case == "group":
            rs = np.random.RandomState(5)
            v = np.round(fixture_array((2, 60), seed=6) * 10) / 2
            labels = rs.randint(-1, 6, size=60)
            tv, tl = torch.from_numpy(v[:, lo:hi].copy()), torch.from_numpy(labels[lo:hi].copy())
            for f in oracle.GROUPED_FUNCS:
                out[f] = nd.group_sharded(f, tv, tl, num_labels=6, ddof=1, index_offset=lo, backend=B).numpy()
        elif case == "reduce":
            x = fixture_array((4, 60), seed=8)
            x[1, :] = np.nan  # all-NaN slice: arg* must raise on every rank
            x[2, 30:] = np.nan
            for f in oracle.AGGREGATION_FUNCS:
                for ax, full in ((-1, x), (0, x.T.copy())):
                    piece = torch.from_numpy((full[:, lo:hi] if ax == -1 else full[lo:hi]).copy())
                    try:
                        out[f"{f}_{ax}"] = nd.reduce_sharded(f, piece, axis=ax, ddof=1, shard_lens=lens, backend=B).numpy()
                    except ValueError as e:
                        out[f"{f}_{ax}"] = str(e)
            y = x[[0, 2, 3]]
            for f in ("nanargmax", "nanargmin"):
                out[f"{f}_ok"] = nd.reduce_sharded(f, torch.from_numpy(y[:, lo:hi].copy()), axis=-1, shard_lens=lens, backend=B).numpy()
        results[rank] = (lo, hi, out)
    finally:
        dist.destroy_process_group()


def _run(case, uneven):
    world = 2
    mgr = mp.Manager()
    results = mgr.dict()
    mp.spawn(_worker, args=(world, _free_port(), case, uneven, results), nprocs=world, join=True)
    return [results[r] for r in range(world)]


def _stitch(res, key, axis=-1):
    return np.concatenate([r[2][key] for r in res], axis=axis)


@pytest.mark.parametrize("uneven", [False, True])
def test_move_sharded_matches_unsharded(uneven):
    res = _run("move", uneven)
    a = fixture_array((3, 60), seed=1)
    b = a**2 + 1
    for f in ("move_mean", "move_sum", "move_std", "move_var"):
        np.testing.assert_allclose(_stitch(res, f), getattr(oracle, f)(a, window=7, min_count=2), rtol=1e-12, equal_nan=True)
    for f in ("move_cov", "move_corr"):
        np.testing.assert_allclose(_stitch(res, f), getattr(oracle, f)(a, b, window=7, min_count=2), rtol=1e-10, equal_nan=True)
    np.testing.assert_allclose(_stitch(res, "move_mean_axis0"), oracle.move_mean(a, window=7, min_count=2), rtol=1e-12, equal_nan=True)


@pytest.mark.parametrize("uneven", [False, True])
def test_move_exp_sharded_matches_unsharded(uneven):
    res = _run("exp", uneven)
    a = fixture_array((3, 60), seed=1)
    b = a**2 + 1
    for f in ("move_exp_nancount", "move_exp_nanmean", "move_exp_nansum", "move_exp_nanvar", "move_exp_nanstd"):
        np.testing.assert_allclose(_stitch(res, f), getattr(oracle, f)(a, alpha=0.2, min_weight=0.1), rtol=1e-11, equal_nan=True)
    for f in ("move_exp_nancov", "move_exp_nancorr"):
        np.testing.assert_allclose(_stitch(res, f), getattr(oracle, f)(a, b, alpha=0.2), rtol=1e-9, equal_nan=True)
    al = np.random.RandomState(3).rand(60) * 0.8 + 0.1
    np.testing.assert_allclose(_stitch(res, "move_exp_nanmean_alpha1d"), oracle.move_exp_nanmean(a, alpha=al), rtol=1e-11, equal_nan=True)


@pytest.mark.parametrize("uneven", [False, True])
def test_move_exp_sharded_single_pass(uneven):
    res = _run("exp_single", uneven)
    x = fixture_array((2, 2400), seed=11)
    x[1, 1100:1300] = np.nan
    for f in ("move_exp_nanmean", "move_exp_nansum", "move_exp_nanvar"):
        np.testing.assert_allclose(_stitch(res, f), getattr(oracle, f)(x, alpha=0.9999), rtol=1e-11, equal_nan=True)


def test_exp_forget_length():
    from numbagg_b200.distributed import exp_forget_length

    import math

    for alpha in (0.1, 0.5, 0.9999, 1e-3):
        k = exp_forget_length(alpha)
        assert k * math.log2(1.0 - alpha) < -1075.0, (alpha, k)  # below half the smallest subnormal
        assert (1.0 - alpha) ** k == 0.0
    assert 7000 < exp_forget_length(0.1) < 7400
    assert exp_forget_length(0.0) is None and exp_forget_length(-0.5) is None and exp_forget_length(float("nan")) is None
    assert exp_forget_length(1.0) == 1


@pytest.mark.parametrize("uneven", [False, True])
def test_fill_sharded_matches_unsharded_exactly(uneven):
    res = _run("fill", uneven)
    x = fixture_array((3, 60), seed=1)
    x[1, 10:50] = np.nan
    x[2, :] = np.nan
    for f in ("ffill", "bfill"):
        for limit in (None, 2, 25):
            np.testing.assert_array_equal(_stitch(res, f"{f}_{limit}"), getattr(oracle, f)(x, limit=limit))


@pytest.mark.parametrize("uneven", [False, True])
def test_group_sharded_matches_unsharded(uneven):
    res = _run("group", uneven)
    rs = np.random.RandomState(5)
    v = np.round(fixture_array((2, 60), seed=6) * 10) / 2
    labels = rs.randint(-1, 6, size=60)
    for f in oracle.GROUPED_FUNCS:
        exp = getattr(oracle, f)(v, labels, num_labels=6, axis=-1)
        for r in res:  # every rank holds the full result
            np.testing.assert_allclose(r[2][f], exp, rtol=1e-12, equal_nan=True, err_msg=f)


@pytest.mark.parametrize("uneven", [False, True])
def test_reduce_sharded_matches_unsharded(uneven):
    res = _run("reduce", uneven)
    x = fixture_array((4, 60), seed=8)
    x[1, :] = np.nan
    x[2, 30:] = np.nan
    for f in oracle.AGGREGATION_FUNCS:
        for r in res:  # every rank holds the full result (or raised the reference's error)
            for ax in (-1, 0):
                got = r[2][f"{f}_{ax}"]
                if f in ("nanargmax", "nanargmin"):
                    assert got == "All-NaN slice encountered"
                    continue
                exp = getattr(oracle, f)(x, axis=-1)
                if exp.dtype.kind == "f":
                    np.testing.assert_allclose(got, exp, rtol=1e-12, equal_nan=True, err_msg=f)
                else:
                    np.testing.assert_array_equal(got, exp, err_msg=f)
    for f in ("nanargmax", "nanargmin"):
        for r in res:
            np.testing.assert_array_equal(r[2][f"{f}_ok"], getattr(oracle, f)(x[[0, 2, 3]], axis=-1))


def test_row_slice_covers_everything():
    from numbagg_b200.distributed import row_slice

    for n in (0, 1, 7, 8, 10_000):
        for world in (1, 2, 3, 8):
            idx = np.concatenate([np.arange(n)[row_slice(n, r, world)] for r in range(world)])
            np.testing.assert_array_equal(idx, np.arange(n))
