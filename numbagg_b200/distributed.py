"""Multi-GPU forms of the hot path: one process per GPU, ``torch.distributed`` for the
exchange (NCCL over NVLink on the GPU box, gloo in the CPU tests), the CUDA kernels for all
arithmetic.  SURVEY.md 8(e):

* rows (non-core slices) shard with ZERO communication -- call the ordinary functions on the
  local rows (``row_slice`` computes the split);
* a long CORE axis shards contiguously in rank order with ONE exchange step and ONE pass over
  the data:
    - moving windows: every rank receives the last `window` elements of its predecessor
      (point-to-point send/recv) WHILE it computes its whole shard without a halo; only the
      first `window` outputs are then recomputed with the halo;
    - move_exp_*: every rank scans its shard from a zero state and reduces it to a per-slice
      aggregate in the same pass; the aggregates (88 bytes per slice) are all-gathered and
      folded, and only the head of the shard -- up to the first position where the decay
      product has underflowed to exactly 0.0, ~7 070 elements for alpha = 0.1 -- is recomputed
      with the carry.  Array alphas / alphas close to 0 take two passes;
    - ffill / bfill: every rank fills its shard with a SENTINEL carry (a NaN payload that no
      output can otherwise hold) and produces its aggregate in the same pass; after the
      all-gather only the leading sentinel run is rewritten (``nbg_fill_patch``);
    - grouped reductions over element shards: per-label partial states are combined with
      collectives on the state's channel planes -- SUM for the additive ops, PRODUCT / MAX /
      MIN for prod / min / max / any / all, and for (value, index) ops a MAX on the
      order-preserving key followed by a MIN on the index masked to the ranks that hold the
      winning key (first / last: MIN / MAX on the index, then the value of the rank that
      holds it).  No rank ever materialises another rank's table;
    - plain reductions over element shards: 3-word state records per output are
      all-gathered and folded by ``nbg_reduce_merge``.

No function here synchronises the host with the device: shard lengths are part of the call
(`shard_lens`), not exchanged.

`backend` (default: the CUDA kernels) exists so that the exchange logic can be exercised by
world_size-2 gloo tests on CPU-only machines; the product itself has no CPU path.
"""

from __future__ import annotations

import math
from typing import Sequence

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
# how each channel of a group state combines across element shards (channel -> reduce op);
# float sums are float64 bit patterns in the int64 state
_GROUP_COMBINE = {
    "group_nansum": {0: "sum_v"}, "group_nansum_of_squares": {0: "sum_v"}, "group_nancount": {2: "sum_i"},
    "group_nanmean": {0: "sum_v", 2: "sum_i"}, "group_nanvar": {0: "sum_v", 1: "sum_v", 2: "sum_i"},
    "group_nanstd": {0: "sum_v", 1: "sum_v", 2: "sum_i"}, "group_nanprod": {0: "prod_v"},
    "group_nanmax": {0: "max_key"}, "group_nanmin": {0: "max_key"}, "group_nanany": {0: "max_i"},
    "group_nanall": {0: "min_i"},
}
_I64_MIN = -(2 ** 63)
_I64_MAX = 2 ** 63 - 1


class CudaBackend:
    """Local compute = the C-ABI kernels (numbagg_b200.decorators.run_*)."""

    move = staticmethod(D.run_move)
    move_exp = staticmethod(D.run_move_exp)
    fill = staticmethod(D.run_fill)
    fill_patch = staticmethod(D.run_fill_patch)
    fill_sentinel = staticmethod(D.fill_sentinel_bits)
    group_partial = staticmethod(D.run_group_partial)
    group_channels = staticmethod(D.group_state_channels)
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


def core_slice(n: int, rank: int, world: int) -> slice:
    """Even contiguous split of a core axis (the convention `shard_lens=None` assumes)."""
    return row_slice(n, rank, world)


def _lens(shard_len: int, shard_lens: Sequence[int] | None, rank: int, world: int) -> list[int]:
    """Per-rank shard lengths along the core axis.  They are host knowledge of the caller that cut
    the array (`shard_lens`); without them every shard is taken to have this rank's length --
    exact for an even split, and never a device round trip."""
    if shard_lens is None:
        return [int(shard_len)] * world
    lens = [int(x) for x in shard_lens]
    if len(lens) != world or lens[rank] != shard_len:
        raise ValueError(f"shard_lens {lens} does not describe this rank's shard (rank {rank}, length {shard_len})")
    return lens


def _all_gather(t: torch.Tensor, group) -> torch.Tensor:
    """(world,) + t.shape: every rank's `t` (same shape everywhere), in rank order."""
    world = dist.get_world_size(group)
    t = t.contiguous()
    out = torch.empty((world * t.shape[0],) + tuple(t.shape[1:]), dtype=t.dtype, device=t.device)
    dist.all_gather_into_tensor(out, t, group=group)
    return out.view((world,) + tuple(t.shape))


def _p2p(group, sends: list[tuple[torch.Tensor, int]], recvs: list[tuple[torch.Tensor, int]]):
    """Batched point-to-point exchange (global ranks resolved from the group)."""
    ops = []
    for t, peer in sends:
        ops.append(dist.P2POp(dist.isend, t, dist.get_global_rank(group, peer) if group is not None else peer, group))
    for t, peer in recvs:
        ops.append(dist.P2POp(dist.irecv, t, dist.get_global_rank(group, peer) if group is not None else peer, group))
    return dist.batch_isend_irecv(ops) if ops else []


# ------------------------------------------------------------------------- moving windows
def move_sharded(name: str, *shards: torch.Tensor, window: int, min_count: int | None = None,
                 axis: int = -1, shard_lens: Sequence[int] | None = None, group=None,
                 backend=CudaBackend) -> torch.Tensor:
    """`shards`: this rank's contiguous piece (along `axis`) of each input, ranks in order.
    Returns this rank's piece of the result.

    Exchange: rank r sends the last min(window, len) elements of its shard to rank r+1 (one
    send/recv pair per input, batched) while the interior is being computed; a shard shorter
    than the window forwards its own halo first (chain), which only happens for degenerate
    splits.  The first `window` outputs are then recomputed from (halo, shard[:window])."""
    rank, world = _world(group)
    if min_count is None:
        min_count = window
    nd = shards[0].dim()
    axis = axis % nd
    n_local = shards[0].shape[axis]
    lens = _lens(n_local, shard_lens, rank, world)
    need = min(window, sum(lens[:rank]))  # elements of history this rank can use
    short_chain = any(ln < window for ln in lens[:-1])
    halos = None
    if short_chain:
        # degenerate split (some shard shorter than the window): relay halos rank by rank
        halos = []
        for s in shards:
            h = None
            if rank > 0 and need > 0:
                shape = list(s.shape)
                shape[axis] = need
                h = torch.empty(shape, dtype=s.dtype, device=s.device)
                for w in _p2p(group, [], [(h, rank - 1)]):
                    w.wait()
            if rank + 1 < world:
                have = torch.cat([h, s], dim=axis) if h is not None else s
                k = min(window, have.shape[axis])
                tail = have.narrow(axis, have.shape[axis] - k, k).contiguous()
                for w in _p2p(group, [(tail, rank + 1)], []):
                    w.wait()
            halos.append(h)
        halos = None if halos[0] is None else halos
        return backend.move(name, list(shards), window, min_count, axis, halos)

    # regular split: post the exchange, compute the whole shard without halo, fix the head
    recv_bufs, works = [], []
    sends, recvs = [], []
    for s in shards:
        if rank + 1 < world:
            sends.append((s.narrow(axis, n_local - window, window).contiguous(), rank + 1))
        if rank > 0:
            shape = list(s.shape)
            shape[axis] = window
            buf = torch.empty(shape, dtype=s.dtype, device=s.device)
            recv_bufs.append(buf)
            recvs.append((buf, rank - 1))
    works = _p2p(group, sends, recvs)
    out = backend.move(name, list(shards), window, min_count, axis, None)
    for w in works:
        w.wait()
    if rank > 0:
        head_len = min(n_local, window)
        heads = [s.narrow(axis, 0, head_len) for s in shards]
        fixed = backend.move(name, heads, window, min_count, axis, recv_bufs)
        k = min(window, n_local)
        out.narrow(axis, 0, k).copy_(fixed.narrow(axis, 0, k))
    return out


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


def exp_fold(name: str, aggs: torch.Tensor, upto: int) -> torch.Tensor | None:
    """Compose aggs[0] .. aggs[upto-1] ((world, slices, NBG_EXP_STATE), gathered in rank order)."""
    carry = None
    for r in range(upto):
        carry = aggs[r] if carry is None else exp_compose(name, carry, aggs[r])
    return carry


def exp_forget_length(alpha: float) -> int | None:
    """Number of steps k after which the decay (1 - alpha)^k is below 2^-1075, i.e. below half the
    smallest subnormal: beyond that position the carry's share `D*carry` of a shard's state
    `s = D*carry + U` cannot be represented any more (and the kernels' D -- a product of chunk and
    tile aggregates -- has underflowed to exactly 0).  Outputs there depend on the carry only
    where the reference's own sums have decayed into the subnormal range (a NaN run longer than
    k positions: DESIGN.md "known parity limits").  None when alpha does not forget (<= 0, >= 2, NaN)."""
    d = abs(1.0 - float(alpha))
    if not (d < 1.0):
        return None
    if d == 0.0:
        return 1
    k = 1075.0 * math.log(2.0) / -math.log(d)
    if not math.isfinite(k) or k > 1e12:
        return None
    return int(k * 1.02) + 64  # 2 % + 64 steps of slack


def move_exp_sharded(name: str, *shards: torch.Tensor, alpha, min_weight: float = 0.0, axis: int = -1,
                     group=None, backend=CudaBackend) -> torch.Tensor:
    """Core-axis sharded move_exp_*.  `alpha`: python float, or this rank's shard of a 1-D /
    N-D alpha tensor.

    Scalar alpha: ONE pass -- scan from the zero state producing outputs and the shard
    aggregate together, all-gather the aggregates, recompute only the first
    `exp_forget_length(alpha)` positions with the folded carry.  Otherwise two passes
    (aggregate, then scan with carry)."""
    rank, world = _world(group)
    nd = shards[0].dim()
    axis = axis % nd
    n_local = shards[0].shape[axis]
    forget = None if D.dev.is_tensor(alpha) else exp_forget_length(alpha)
    single_pass = forget is not None and forget * 4 <= n_local
    if single_pass:
        out, agg = backend.move_exp(name, list(shards), alpha, min_weight, axis, None, True, True)
    else:
        out, agg = None, backend.move_exp(name, list(shards), alpha, min_weight, axis, None, True, False)[1]
    aggs = _all_gather(agg, group)
    carry = exp_fold(name, aggs, rank)
    if not single_pass:
        return backend.move_exp(name, list(shards), alpha, min_weight, axis,
                                carry.contiguous() if carry is not None else None, False, True)[0]
    if carry is not None:
        heads = [s.narrow(axis, 0, forget) for s in shards]
        fixed = backend.move_exp(name, heads, alpha, min_weight, axis, carry.contiguous(), False, True)[0]
        out.narrow(axis, 0, forget).copy_(fixed)
    return out


# --------------------------------------------------------------------------------- fills
def fill_compose(older: torch.Tensor, newer: torch.Tensor) -> torch.Tensor:
    """(slices, 3) int64 [has_valid, value bits, distance]: `newer` hides `older` whenever it
    holds a valid value, otherwise the distance keeps growing (nbg_fill.cu)."""
    has_new = (newer[:, 0] != 0).unsqueeze(1)
    grown = torch.stack([older[:, 0], older[:, 1], older[:, 2] + newer[:, 2]], dim=1)
    return torch.where(has_new, newer, grown)


def fill_sharded(name: str, shard: torch.Tensor, *, limit: int | None = None, axis: int = -1,
                 total_len: int | None = None, shard_lens: Sequence[int] | None = None, group=None,
                 backend=CudaBackend) -> torch.Tensor:
    """Core-axis sharded ffill / bfill in ONE pass over the shard.  `limit=None` means the FULL
    axis length (the reference's default, decorators.py:474-475): the sum of `shard_lens`, or
    `total_len`, or world * this shard's length (even split)."""
    rank, world = _world(group)
    nd = shard.dim()
    axis = axis % nd
    n_local = shard.shape[axis]
    if limit is None:
        if total_len is None:
            total_len = sum(_lens(n_local, shard_lens, rank, world))
        limit = total_len
    view_inner = math.prod(shard.shape[axis + 1:])
    slices = shard.numel() // max(n_local, 1)
    sentinel = backend.fill_sentinel(shard.element_size())
    single_pass = view_inner == 1 and shard.is_contiguous() and n_local > 0
    if single_pass:
        seed = torch.tensor([1, sentinel - (1 << 64) if sentinel >= (1 << 63) else sentinel, 0],
                            dtype=torch.int64, device=shard.device).repeat(slices, 1)
        out, agg = backend.fill(name, shard, limit, axis, seed, True, True)
        # an all-NaN shard reports the sentinel itself as its last value: it holds no value
        empty = agg[:, 1] == seed[0, 1]
        agg = torch.where(empty.unsqueeze(1), torch.stack([torch.zeros_like(agg[:, 0]), torch.zeros_like(agg[:, 1]),
                                                          torch.full_like(agg[:, 2], n_local)], dim=1), agg)
    else:
        out, agg = None, backend.fill(name, shard, limit, axis, None, True, False)[1]
    aggs = _all_gather(agg, group)
    # ffill: carry comes from lower ranks in ascending order; bfill scans from the far end
    order = range(rank) if name == "ffill" else range(world - 1, rank, -1)
    carry = None
    for r in order:
        carry = aggs[r] if carry is None else fill_compose(carry, aggs[r])
    if not single_pass:
        return backend.fill(name, shard, limit, axis, carry.contiguous() if carry is not None else None, False, True)[0]
    if carry is None:
        carry = torch.zeros((slices, _lib.NBG_FILL_STATE), dtype=torch.int64, device=shard.device)
    backend.fill_patch(name, out, limit, axis, carry.contiguous())
    return out


# ------------------------------------------------------------------------------- grouped
def _sign_flip(t: torch.Tensor) -> torch.Tensor:
    """u64 order (stored in int64 words) -> int64 order, and back (an involution)."""
    return t ^ _I64_MIN


def group_sharded(name: str, values: torch.Tensor, labels: torch.Tensor, *, num_labels: int, ddof: int = 1,
                  index_offset: int = 0, labels_per_row: bool = False, group=None,
                  backend=CudaBackend) -> torch.Tensor:
    """values (rows, n_local), labels (n_local,) or (rows, n_local): this rank's ELEMENT shard
    (contiguous along the core axis, ranks in order; `index_offset` = flat index of its first
    element, used by arg*/first/last).  Every rank returns the full (rows, num_labels) result.

    Exchange: all-reduces on the channel planes of the partial state (see the module
    docstring); bytes on the wire per rank are those of ONE table, whatever the world size."""
    rank, world = _world(group)
    vdtype = D.dev.np_dtype_of(values)
    state = backend.group_partial(name, values, labels, num_labels, index_offset, labels_per_row)
    ch = backend.group_channels(name, state, values.shape[0], num_labels)
    is_float = vdtype.kind == "f"

    def reduce_plane(c: int, how: str):
        plane = ch[c]
        dense = plane.contiguous()
        if how in ("sum_v", "prod_v") and is_float:
            work = dense.view(torch.float64)
            dist.all_reduce(work, op=dist.ReduceOp.SUM if how == "sum_v" else dist.ReduceOp.PRODUCT, group=group)
            dense = work.view(torch.int64)
        elif how in ("sum_v", "sum_i"):
            dist.all_reduce(dense, op=dist.ReduceOp.SUM, group=group)
        elif how == "prod_v":
            dist.all_reduce(dense, op=dist.ReduceOp.PRODUCT, group=group)
        elif how == "max_key":
            dense = _sign_flip(dense)
            dist.all_reduce(dense, op=dist.ReduceOp.MAX, group=group)
            dense = _sign_flip(dense)
        elif how == "max_i":
            dist.all_reduce(dense, op=dist.ReduceOp.MAX, group=group)
        elif how == "min_i":
            dist.all_reduce(dense, op=dist.ReduceOp.MIN, group=group)
        else:
            raise AssertionError(how)
        plane.copy_(dense)

    if name in _GROUP_COMBINE:
        for c, how in _GROUP_COMBINE[name].items():
            reduce_plane(c, how)
    elif name in ("group_nanargmax", "group_nanargmin"):
        # ch0 = order-preserving key of the extreme (0 = empty), ch1 = global index of its first
        # occurrence: MAX on the key, then MIN on the index among the ranks holding that key
        key = _sign_flip(ch[0].contiguous())
        best = key.clone()
        dist.all_reduce(best, op=dist.ReduceOp.MAX, group=group)
        idx = torch.where(key == best, ch[1].contiguous(), torch.full_like(key, _I64_MAX))
        dist.all_reduce(idx, op=dist.ReduceOp.MIN, group=group)
        ch[0].copy_(_sign_flip(best))
        ch[1].copy_(idx)
    elif name in ("group_nanfirst", "group_nanlast"):
        # ch1 = global index of the first / last valid element (INT64_MAX / -1 = none), ch0 = its bits:
        # MIN / MAX on the index, then the bits of the one rank that holds it (others contribute 0)
        idx = ch[1].contiguous()
        win = idx.clone()
        dist.all_reduce(win, op=dist.ReduceOp.MIN if name == "group_nanfirst" else dist.ReduceOp.MAX, group=group)
        none = _I64_MAX if name == "group_nanfirst" else -1
        bits = torch.where((idx == win) & (win != none), ch[0].contiguous(), torch.zeros_like(idx))
        dist.all_reduce(bits, op=dist.ReduceOp.SUM, group=group)
        ch[0].copy_(bits)
        ch[1].copy_(win)
    else:
        raise ValueError(f"unknown grouped function {name}")
    return backend.group_finalize(name, vdtype, state, ddof)


def reduce_sharded(name: str, shard: torch.Tensor, *, axis: int = -1, ddof: int = 1,
                   shard_lens: Sequence[int] | None = None, group=None, backend=CudaBackend) -> torch.Tensor:
    """Plain NaN reduction (allnan ... nanmin) of an array sharded along the ONE reduced axis
    `axis`: `shard` is this rank's contiguous piece, ranks in order (`shard_lens` when the
    pieces differ in length).  One exchange: the per-output state records (3 words each) are
    all-gathered and every rank folds them, so every rank returns the full result.  nanarg*
    return positions in the unsharded axis and raise like the reference on an all-NaN slice (the
    only host round trip in this module); empty reductions raise like the reference."""
    rank, world = _world(group)
    axis %= shard.dim()
    work = D._reduce_loop_dtype(name, D.dev.np_dtype_of(shard))
    tdt = D.dev._NP_TO_TORCH[work]
    if shard.dtype != tdt:
        shard = shard.to(tdt)
    lens = _lens(shard.shape[axis], shard_lens, rank, world)
    n_total = sum(lens)
    batch = [s for d, s in enumerate(shard.shape) if d != axis]
    if n_total == 0 and int(np.prod(batch)) > 0 and name in D._REDUCE_EMPTY_ERRORS:
        raise ValueError(D._REDUCE_EMPTY_ERRORS[name])
    states, restore = backend.reduce_partial(name, shard, (axis,), sum(lens[:rank]))
    gathered = _all_gather(states, group)
    out = restore(backend.reduce_merge(name, work, gathered, n_total, ddof))
    if name in ("nanargmax", "nanargmin") and bool((out < 0).any()):  # the one host round trip: the reference raises
        raise ValueError("All-NaN slice encountered")
    return out
