"""The sharded forms of the hot path over REAL NCCL with the CUDA kernels, torchrun-free (one
spawned process per GPU): used by tests/test_distributed_gpu.py and by __graft_entry__.smoke()
whenever >= 2 GPUs are visible.  Checks every sharded result against the oracle (the CPU
restatement of the reference; checker only)."""

import os
import socket

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def fixture_array(shape, nan_frac=0.2, seed=0):
    a = np.random.RandomState(seed).rand(*shape)
    return np.where(a > nan_frac, a, np.nan)


N = 400_000
K = 5000
GROUP_FUNCS = ("group_nanargmax", "group_nanargmin", "group_nanfirst", "group_nanlast", "group_nanvar", "group_nansum",
               "group_nanmean", "group_nanmax", "group_nanmin", "group_nanprod", "group_nancount", "group_nanany", "group_nanall")
REDUCE_FUNCS = ("nansum", "nanvar", "nanargmax", "nanmax", "nancount", "anynan")


def _inputs():
    a = fixture_array((3, N), seed=1)
    a[1, 150_000:260_000] = np.nan  # a NaN run across the 2-rank boundary (and longer than exp's memory)
    a[2, :10] = np.nan
    b = a**2 + 1
    labels = np.random.RandomState(5).randint(-1, K, size=N)
    # exponential functions: a NaN run across the boundary, but shorter than the ~7 070 steps after which
    # the reference's decayed sums turn subnormal (DESIGN.md "known parity limits")
    ae = fixture_array((3, N), seed=1)
    ae[1, 198_500:201_500] = np.nan
    return a, b, labels, ae


def _worker(rank, world, port, results):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    try:
        from numbagg_b200 import distributed as nd

        a, b, labels, ae = _inputs()
        lo, hi = N * rank // world, N * (rank + 1) // world
        lens = [N * (r + 1) // world - N * r // world for r in range(world)]
        ta = torch.from_numpy(a[:, lo:hi].copy()).cuda()
        tb = torch.from_numpy(b[:, lo:hi].copy()).cuda()
        te = torch.from_numpy(ae[:, lo:hi].copy()).cuda()
        te2 = te * te + 1
        out = {}
        out["move_std"] = nd.move_sharded("move_std", ta, window=1000, min_count=500, shard_lens=lens).cpu().numpy()
        out["move_corr"] = nd.move_sharded("move_corr", ta, tb, window=1000, min_count=500, shard_lens=lens).cpu().numpy()
        out["move_mean_w7"] = nd.move_sharded("move_mean", ta, window=7, min_count=1, shard_lens=lens).cpu().numpy()
        out["move_exp_nanmean"] = nd.move_exp_sharded("move_exp_nanmean", te, alpha=0.1).cpu().numpy()          # one pass
        out["move_exp_nanvar"] = nd.move_exp_sharded("move_exp_nanvar", te, alpha=0.1).cpu().numpy()            # one pass
        out["move_exp_nanmean_slow"] = nd.move_exp_sharded("move_exp_nanmean", te, alpha=1e-4).cpu().numpy()    # two passes
        out["move_exp_nancorr"] = nd.move_exp_sharded("move_exp_nancorr", te, te2, alpha=0.3).cpu().numpy()
        out["ffill"] = nd.fill_sharded("ffill", ta, shard_lens=lens).cpu().numpy()
        out["bfill"] = nd.fill_sharded("bfill", ta, limit=3, shard_lens=lens).cpu().numpy()
        out["ffill_limit"] = nd.fill_sharded("ffill", ta, limit=60_000, shard_lens=lens).cpu().numpy()
        out["bfill_all"] = nd.fill_sharded("bfill", ta, shard_lens=lens).cpu().numpy()
        out["ffill_f32"] = nd.fill_sharded("ffill", ta.float(), shard_lens=lens).cpu().numpy()
        tl = torch.from_numpy(labels[lo:hi].copy()).cuda()
        va = torch.from_numpy(np.round(a[:, lo:hi] * 20) / 4).cuda()  # many ties for arg* / first / last
        for f in GROUP_FUNCS:
            vv = (1.0 + va / 64) if f == "group_nanprod" else va
            out[f] = nd.group_sharded(f, vv, tl, num_labels=K, index_offset=lo).cpu().numpy()
        for f in REDUCE_FUNCS:
            out[f] = nd.reduce_sharded(f, ta[[0, 2]], axis=-1, shard_lens=lens).cpu().numpy()
        results[rank] = out
    finally:
        dist.destroy_process_group()


def run(world: int = 2) -> dict:
    """Spawn `world` ranks, run every sharded form, compare with the oracle.  Returns a summary
    (raises AssertionError on any mismatch)."""
    from oracle import oracle

    mgr = mp.Manager()
    results = mgr.dict()
    mp.spawn(_worker, args=(world, _free_port(), results), nprocs=world, join=True)
    a, b, labels, ae = _inputs()
    cat = lambda k: np.concatenate([results[r][k] for r in range(world)], axis=1)  # noqa: E731
    close = lambda g, e, rt: (np.testing.assert_array_equal(np.isnan(g), np.isnan(e)),  # noqa: E731
                              np.testing.assert_allclose(g, e, rtol=rt, atol=rt * 1e-3, equal_nan=True))
    # variances are differences of window sums: compare them as variances with a floor on that scale
    np.testing.assert_allclose(cat("move_std") ** 2, oracle.move_std(a, window=1000, min_count=500) ** 2, rtol=1e-9, atol=1e-12, equal_nan=True)
    np.testing.assert_allclose(cat("move_corr"), oracle.move_corr(a, b, window=1000, min_count=500), rtol=1e-8, atol=1e-9, equal_nan=True)
    close(cat("move_mean_w7"), oracle.move_mean(a, window=7, min_count=1), 1e-12)
    close(cat("move_exp_nanmean"), oracle.move_exp_nanmean(ae, alpha=0.1), 1e-12)
    close(cat("move_exp_nanvar"), oracle.move_exp_nanvar(ae, alpha=0.1), 1e-9)
    close(cat("move_exp_nanmean_slow"), oracle.move_exp_nanmean(ae, alpha=1e-4), 1e-11)
    close(cat("move_exp_nancorr"), oracle.move_exp_nancorr(ae, ae**2 + 1, alpha=0.3), 1e-8)
    np.testing.assert_array_equal(cat("ffill"), oracle.ffill(a))
    np.testing.assert_array_equal(cat("bfill"), oracle.bfill(a, limit=3))
    np.testing.assert_array_equal(cat("ffill_limit"), oracle.ffill(a, limit=60_000))
    np.testing.assert_array_equal(cat("bfill_all"), oracle.bfill(a))
    np.testing.assert_array_equal(cat("ffill_f32"), oracle.ffill(a.astype(np.float32)))
    va = np.round(a * 20) / 4
    for f in GROUP_FUNCS:
        vv = (1.0 + va / 64) if f == "group_nanprod" else va
        exp = getattr(oracle, f)(vv, labels, num_labels=K, axis=-1)
        for r in range(world):  # every rank holds the full result
            if f in ("group_nanvar", "group_nansum", "group_nanmean", "group_nanprod"):
                np.testing.assert_allclose(results[r][f], exp, rtol=1e-11, atol=1e-12, equal_nan=True, err_msg=f)
            else:
                np.testing.assert_array_equal(results[r][f], exp, err_msg=f)
    for f in REDUCE_FUNCS:
        exp = getattr(oracle, f)(a[[0, 2]], axis=-1)
        for r in range(world):
            if f in ("nansum", "nanvar"):
                np.testing.assert_allclose(results[r][f], exp, rtol=1e-12, err_msg=f)
            else:
                np.testing.assert_array_equal(results[r][f], exp, err_msg=f)
    return dict(world=world, checks=12 + len(GROUP_FUNCS) + len(REDUCE_FUNCS))
