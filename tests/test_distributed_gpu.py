"""NCCL version of tests/test_distributed_cpu.py with the real CUDA kernels: needs >= 2 GPUs
(skipped otherwise).  One process per GPU, torch.distributed over NCCL (tests/_nccl_check.py,
which __graft_entry__.smoke() also runs when it sees two GPUs)."""

import pytest
import torch

pytestmark = pytest.mark.gpu


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs >= 2 GPUs")
def test_sharded_matches_unsharded_nccl():
    from tests import _nccl_check

    summary = _nccl_check.run(2)
    assert summary["checks"] > 20


@pytest.mark.skipif(torch.cuda.device_count() < 4, reason="needs >= 4 GPUs")
def test_sharded_matches_unsharded_nccl_4():
    from tests import _nccl_check

    _nccl_check.run(4)


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs >= 2 GPUs")
def test_tensor_on_other_device_runs_there():
    """a tensor on cuda:1 while cuda:0 is current must be processed on cuda:1 (pointers,
    workspaces and stream of ONE device); operands on two devices are an error."""
    import numpy as np

    import numbagg_b200 as nb
    from oracle import oracle

    torch.cuda.set_device(0)
    a = np.random.RandomState(0).rand(50, 3000)
    a[a < 0.2] = np.nan
    t1 = torch.from_numpy(a).to("cuda:1")
    got = nb.move_mean(t1, window=20, min_count=1)
    assert got.device == t1.device and torch.cuda.current_device() == 0
    np.testing.assert_allclose(got.cpu().numpy(), oracle.move_mean(a, window=20, min_count=1), rtol=1e-12, equal_nan=True)
    labels = torch.from_numpy(np.random.RandomState(1).randint(0, 7, size=3000)).to("cuda:1")
    g = nb.group_nansum(t1, labels, num_labels=7, axis=-1)
    np.testing.assert_allclose(g.cpu().numpy(), oracle.group_nansum(a, labels.cpu().numpy(), num_labels=7, axis=-1), rtol=1e-12)
    np.testing.assert_array_equal(nb.ffill(t1).cpu().numpy(), oracle.ffill(a))
    with pytest.raises(ValueError, match="different devices"):
        nb.move_cov(t1, t1.to("cuda:0"), window=5)
