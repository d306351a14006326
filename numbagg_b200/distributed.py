"""Multi-GPU forms of the hot path: one process per GPU, ``torch.distributed`` for the
exchange (NCCL over NVLink on the GPU box, gloo in the CPU tests), the CUDA kernels for all
arithmetic.  SURVEY.md 8(e):

* rows (non-core slices) shard with ZERO communication -- call the ordinary functions on the
  local rows (``row_slice`` computes the split);
* a long CORE axis shards contiguously in rank order with ONE exchange step:
    - moving windows: every rank needs the `window` elements preceding its shard (halo);
    - ffill/bfill and move_exp_*: every rank reduces its shard to a per-slice aggregate,
      the aggregates are all-gathered, each rank folds its predecessors' into a carry and
      then scans its shard with that carry (``carry_in`` of the C ABI);
    - grouped reductions over element shards: per-label partial states are combined --
      ``all_reduce(SUM)`` for the additive ops, all-gather + ordered merge for the rest;
    - plain reductions over element shards: 3-word state records per output are
      all-gathered and folded by ``nbg_reduce_merge``.

`backend` (default: the CUDA kernels) exists so that the exchange logic can be exercised by
world_size-2 gloo tests on CPU-only machines; the product itself has no CPU path.
"""

from __future__ import annotations

import numpy as np
import torch
import torch.distributed as dist

from . import _lib
from . import decorators as D

# channel that decays with d^2 per op (nbg_move_exp.cu), index into state[2:10]
_EXP_SQ_CHANNEL = {
    "move_exp_nancount": None, "move_exp_nanmean": None, "move_exp_nansum": None,
    "move_exp_nanvar": 3, "move_exp_nanstd": 3, "move_exp_nancov": 4, "move_exp_nancorr": 4,
}
# additive ops: (float-sum slots, count slots) inside a record (ws_layout() in nbg_group.cu)
_ADDITIVE_GROUP_OPS = {
    "group_nansum": ([0], []), "group_nansum_of_squares": ([0], []), "group_nancount": ([], [0]),
    "group_nanmean": ([0], [1]), "group_nanvar": ([0, 1], [2]), "group_nanstd": ([0, 1], [2]),
}


class CudaBackend:
    """Local compute = the C-ABI kernels (numbagg_b200.decorators.run_*)."""

    move = staticmethod(D.run_move)
    move_exp = staticmethod(D.run_move_exp)
    fill = staticmethod(D.run_fill)
    group_partial = staticmethod(D.run_group_partial)
    group_combine = staticmethod(D.run_group_combine)
    group_finalize = staticmethod(D.run_group_finalize)
    reduce_merge = staticmethod(D.run_reduce_merge)

    @staticmethod
    def reduce_partial(name, shard, axes, index_offset):
        states, view = D.run_reduce_partial(name, shard, axes, index_offset)
        return states, view.restore


def _world(group):
    return dist.get_rank(group), dist.get_world_size(group)


def row_slice(n_rows: int, rank: int, world: int) -> slice:
    """Even contiguous split of independent rows; no communication is ever needed."""
    base, rem = divmod(n_rows, world)
    start = rank * base + min(rank, rem)
    return slice(start, start + base + (1 if rank < rem else 0))


def _all_gather_var(t: torch.Tensor, axis: int, group) -> list[torch.Tensor]:
    """all_gather of tensors whose length along `axis` may differ between ranks."""
    rank, world = _world(group)
    n_local = torch.tensor([t.shape[axis]], dtype=torch.int64, device=t.device)
    lens = [torch.zeros_like(n_local) for _ in range(world)]
    dist.all_gather(lens, n_local, group=group)
    lens = [int(x.item()) for x in lens]
    m = max(lens)
    moved = t.movedim(axis, 0).contiguous()
    pad = torch.zeros((m,) + tuple(moved.shape[1:]), dtype=t.dtype, device=t.device)
    pad[: moved.shape[0]] = moved
    bufs = [torch.empty_like(pad) for _ in range(world)]
    dist.all_gather(bufs, pad, group=group)
    return [b[:ln].movedim(0, axis) for b, ln in zip(bufs, lens)]


# ------------------------------------------------------------------------- moving windows
def move_sharded(name: str, *shards: torch.Tensor, window: int, min_count: int | None = None,
                 axis: int = -1, group=None, backend=CudaBackend) -> torch.Tensor:
    """`shards`: this rank's contiguous piece (along `axis`) of each input, ranks in order.
    Returns this rank's piece of the result.  Exchange: the last min(window, len) elements
    of every shard are all-gathered (rows x window elements per rank -- a few MB at most)
    and each rank assembles the `window` elements that precede it."""
    rank, world = _world(group)
    if min_count is None:
        min_count = window
    axis = axis % shards[0].dim()
    halos = []
    for s in shards:
        n_local = s.shape[axis]
        tail = s.narrow(axis, max(0, n_local - window), min(window, n_local))
        tails = _all_gather_var(tail, axis, group)
        prev = tails[:rank]
        if prev:
            h = torch.cat(prev, dim=axis)
            if h.shape[axis] > window:
                h = h.narrow(axis, h.shape[axis] - window, window)
            halos.append(h.contiguous())
    if not halos or halos[0].shape[axis] == 0:
        halos = None
    return backend.move(name, list(shards), window, min_count, axis, halos)


# -------------------------------------------------------------------- exponential moving
def exp_compose(name: str, older: torch.Tensor, newer: torch.Tensor) -> torch.Tensor:
    """(slices, NBG_EXP_STATE) aggregates: state after `older` then `newer`
    (s -> D*s + U per channel, D^2 for the squared-weight channel; nbg_move_exp.cu)."""
    out = torch.empty_like(older)
    out[:, 0] = older[:, 0] * newer[:, 0]
    out[:, 1] = older[:, 1] * newer[:, 1]
    decay = newer[:, 0:1].expand(-1, 8).clone()
    sq = _EXP_SQ_CHANNEL[name]
    if sq is not None:
        decay[:, sq] = newer[:, 1]
    out[:, 2:10] = decay * older[:, 2:10] + newer[:, 2:10]
    out[:, 10] = torch.maximum(older[:, 10], newer[:, 10])
    return out


def move_exp_sharded(name: str, *shards: torch.Tensor, alpha, min_weight: float = 0.0, axis: int = -1,
                     group=None, backend=CudaBackend) -> torch.Tensor:
    """Core-axis sharded move_exp_*.  `alpha`: python float, or this rank's shard of a 1-D /
    N-D alpha tensor.  Two passes over the local shard (aggregate, then scan with carry)."""
    rank, world = _world(group)
    _, agg = backend.move_exp(name, list(shards), alpha, min_weight, axis, None, True, False)
    aggs = [torch.empty_like(agg) for _ in range(world)]
    dist.all_gather(aggs, agg.contiguous(), group=group)
    carry = None
    for r in range(rank):
        carry = aggs[r] if carry is None else exp_compose(name, carry, aggs[r])
    out, _ = backend.move_exp(name, list(shards), alpha, min_weight, axis,
                              carry.contiguous() if carry is not None else None, False, True)
    return out


# --------------------------------------------------------------------------------- fills
def fill_compose(older: torch.Tensor, newer: torch.Tensor) -> torch.Tensor:
    """(slices, 3) int64 [has_valid, value bits, distance]: `newer` hides `older` whenever it
    holds a valid value, otherwise the distance keeps growing (nbg_fill.cu)."""
    has_new = newer[:, 0] != 0
    out = older.clone()
    out[:, 2] = older[:, 2] + newer[:, 2]
    out[has_new] = newer[has_new]
    return out


def fill_sharded(name: str, shard: torch.Tensor, *, limit: int | None = None, axis: int = -1,
                 total_len: int | None = None, group=None, backend=CudaBackend) -> torch.Tensor:
    """Core-axis sharded ffill / bfill.  `limit=None` means the FULL axis length (the
    reference's default, decorators.py:474-475): pass `total_len` or it is all-reduced."""
    rank, world = _world(group)
    if limit is None:
        if total_len is None:
            t = torch.tensor([shard.shape[axis]], dtype=torch.int64, device=shard.device)
            dist.all_reduce(t, group=group)
            total_len = int(t.item())
        limit = total_len
    _, agg = backend.fill(name, shard, limit, axis, None, True, False)
    aggs = [torch.empty_like(agg) for _ in range(world)]
    dist.all_gather(aggs, agg.contiguous(), group=group)
    # ffill: carry comes from lower ranks in ascending order; bfill scans from the far end
    order = range(rank) if name == "ffill" else range(world - 1, rank, -1)
    carry = None
    for r in order:
        carry = aggs[r] if carry is None else fill_compose(carry, aggs[r])
    out, _ = backend.fill(name, shard, limit, axis, carry.contiguous() if carry is not None else None, False, True)
    return out


# ------------------------------------------------------------------------------- grouped
def group_sharded(name: str, values: torch.Tensor, labels: torch.Tensor, *, num_labels: int, ddof: int = 1,
                  index_offset: int = 0, labels_per_row: bool = False, group=None,
                  backend=CudaBackend) -> torch.Tensor:
    """values (rows, n_local), labels (n_local,) or (rows, n_local): this rank's ELEMENT shard
    (contiguous along the core axis, ranks in order; `index_offset` = flat index of its first
    element, used by arg*/first/last).  Every rank returns the full (rows, num_labels) result.
    Exchange: all_reduce(SUM) of the partial states for additive ops; otherwise all-gather of
    the states and an ordered merge (later shards never override earlier ties)."""
    rank, world = _world(group)
    vdtype = D.dev.np_dtype_of(values)
    state = backend.group_partial(name, values, labels, num_labels, index_offset, labels_per_row)
    if name in _ADDITIVE_GROUP_OPS:
        # sums of float data are float64 bit patterns: reduce them through a float64 view
        sum_slots, count_slots = _ADDITIVE_GROUP_OPS[name]
        if vdtype.kind == "f":
            for sl in sum_slots:
                t = state[..., sl].contiguous().view(torch.float64)
                dist.all_reduce(t, group=group)
                state[..., sl] = t.view(torch.int64)
            for sl in count_slots:
                t = state[..., sl].contiguous()
                dist.all_reduce(t, group=group)
                state[..., sl] = t
        else:
            dist.all_reduce(state, group=group)
        total = state
    else:
        states = [torch.empty_like(state) for _ in range(world)]
        dist.all_gather(states, state.contiguous(), group=group)
        total = states[0].clone()
        for r in range(1, world):
            backend.group_combine(name, vdtype, total, states[r])
    return backend.group_finalize(name, vdtype, total, ddof)


def reduce_sharded(name: str, shard: torch.Tensor, *, axis: int = -1, ddof: int = 1, group=None,
                   backend=CudaBackend) -> torch.Tensor:
    """Plain NaN reduction (allnan ... nanmin) of an array sharded along the ONE reduced axis
    `axis`: `shard` is this rank's contiguous piece, ranks in order (pieces may differ in
    length).  One exchange: the per-output state records (3 words each) are all-gathered and
    every rank folds them, so every rank returns the full result.  nanarg* return positions in
    the unsharded axis.  Raises like the reference on all-NaN / empty slices."""
    rank, world = _world(group)
    axis %= shard.dim()
    vdtype = D.dev.np_dtype_of(shard)
    lens = [torch.zeros(1, dtype=torch.int64, device=shard.device) for _ in range(world)]
    dist.all_gather(lens, torch.tensor([shard.shape[axis]], dtype=torch.int64, device=shard.device), group=group)
    lens = [int(x.item()) for x in lens]
    n_total = sum(lens)
    batch = [s for d, s in enumerate(shard.shape) if d != axis]
    if n_total == 0 and int(np.prod(batch)) > 0 and name in D._REDUCE_EMPTY_ERRORS:
        raise ValueError(D._REDUCE_EMPTY_ERRORS[name])
    states, restore = backend.reduce_partial(name, shard, (axis,), sum(lens[:rank]))
    gathered = [torch.empty_like(states) for _ in range(world)]
    dist.all_gather(gathered, states.contiguous(), group=group)
    out = restore(backend.reduce_merge(name, vdtype, torch.stack(gathered), n_total, ddof))
    if name in ("nanargmax", "nanargmin") and bool((out < 0).any()):
        raise ValueError("All-NaN slice encountered")
    return out
