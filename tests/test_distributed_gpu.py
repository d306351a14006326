"""NCCL version of tests/test_distributed_cpu.py with the real CUDA kernels: needs >= 2 GPUs
(skipped otherwise).  One process per GPU, torch.distributed over NCCL."""

import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from oracle import oracle

pytestmark = pytest.mark.gpu


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def fixture_array(shape, nan_frac=0.2, seed=0):
    a = np.random.RandomState(seed).rand(*shape)
    return np.where(a > nan_frac, a, np.nan)


def _worker(rank, world, port, results):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    try:
        from numbagg_b200 import distributed as nd

        n = 400_000
        a = fixture_array((3, n), seed=1)
        b = a**2 + 1
        lo, hi = n * rank // world, n * (rank + 1) // world
        ta = torch.from_numpy(a[:, lo:hi].copy()).cuda()
        tb = torch.from_numpy(b[:, lo:hi].copy()).cuda()
        out = {}
        out["move_std"] = nd.move_sharded("move_std", ta, window=1000, min_count=500).cpu().numpy()
        out["move_corr"] = nd.move_sharded("move_corr", ta, tb, window=1000, min_count=500).cpu().numpy()
        out["move_exp_nanmean"] = nd.move_exp_sharded("move_exp_nanmean", ta, alpha=0.1).cpu().numpy()
        out["ffill"] = nd.fill_sharded("ffill", ta).cpu().numpy()
        out["bfill"] = nd.fill_sharded("bfill", ta, limit=3).cpu().numpy()
        labels = np.random.RandomState(5).randint(0, 5000, size=n)
        tl = torch.from_numpy(labels[lo:hi].copy()).cuda()
        for f in ("group_nanargmax", "group_nanfirst", "group_nanvar", "group_nansum", "group_nanlast"):
            out[f] = nd.group_sharded(f, ta, tl, num_labels=5000, index_offset=lo).cpu().numpy()
        for f in ("nansum", "nanvar", "nanargmax", "nanmax", "nancount", "anynan"):
            out[f] = nd.reduce_sharded(f, ta, axis=-1).cpu().numpy()
        results[rank] = out
    finally:
        dist.destroy_process_group()


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs >= 2 GPUs")
def test_sharded_matches_unsharded_nccl():
    world = 2
    mgr = mp.Manager()
    results = mgr.dict()
    mp.spawn(_worker, args=(world, _free_port(), results), nprocs=world, join=True)
    n = 400_000
    a = fixture_array((3, n), seed=1)
    b = a**2 + 1
    cat = lambda k: np.concatenate([results[r][k] for r in range(world)], axis=1)
    np.testing.assert_allclose(cat("move_std"), oracle.move_std(a, window=1000, min_count=500), rtol=1e-9, equal_nan=True)
    np.testing.assert_allclose(cat("move_corr"), oracle.move_corr(a, b, window=1000, min_count=500), rtol=1e-9, equal_nan=True)
    np.testing.assert_allclose(cat("move_exp_nanmean"), oracle.move_exp_nanmean(a, alpha=0.1), rtol=1e-12, equal_nan=True)
    np.testing.assert_array_equal(cat("ffill"), oracle.ffill(a))
    np.testing.assert_array_equal(cat("bfill"), oracle.bfill(a, limit=3))
    labels = np.random.RandomState(5).randint(0, 5000, size=n)
    for f in ("group_nanargmax", "group_nanfirst", "group_nanlast"):
        np.testing.assert_array_equal(results[0][f], getattr(oracle, f)(a, labels, num_labels=5000, axis=-1))
    for f in ("nanargmax", "nanmax", "nancount", "anynan"):
        np.testing.assert_array_equal(results[1][f], getattr(oracle, f)(a, axis=-1))
    for f in ("nansum", "nanvar"):
        np.testing.assert_allclose(results[0][f], getattr(oracle, f)(a, axis=-1), rtol=1e-12)
    for f in ("group_nanvar", "group_nansum"):
        np.testing.assert_allclose(results[1][f], getattr(oracle, f)(a, labels, num_labels=5000, axis=-1), rtol=1e-11, equal_nan=True)
