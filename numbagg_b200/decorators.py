"""Host-side mirror of numbagg's dispatch decorators for the hot path.

Same class names, call signatures (keyword-only scalars), defaults, validation order and
exception types/messages as numbagg/decorators.py -- ``ndmove`` (:275-341), ``ndmoveexp``
(:344-414), ``ndfill`` (:417-487), ``groupndreduce`` (:490-674) -- but instead of building a
Numba gufunc each ``__call__`` hands device pointers to libnbg_b200.so (include/nbg_b200.h).
There is no CPU path: numpy inputs are copied to the GPU and the result is copied back;
CUDA ``torch.Tensor`` inputs are used in place and a CUDA tensor is returned.
"""

from __future__ import annotations

import inspect
import logging
import math
from typing import Any, Callable

import numpy as np
import torch

from . import _device as dev
from . import _lib

logger = logging.getLogger(__name__)

_F32 = np.dtype(np.float32)
_F64 = np.dtype(np.float64)
_NBG_DTYPE = {
    _F32: _lib.NBG_F32,
    _F64: _lib.NBG_F64,
    np.dtype(np.int32): _lib.NBG_I32,
    np.dtype(np.int64): _lib.NBG_I64,
}


def _float_loop_dtype(*dtypes: np.dtype) -> np.dtype:
    """Loop selection NumPy performs over the (float32, float64) gufunc loops the reference
    registers: the FIRST loop every operand casts to safely -- float16, bool and
    8/16-bit integers reach float32, everything else float64 (probed against the reference:
    move_mean(int8) is float32, move_mean(int32) float64)."""
    dt = np.result_type(*dtypes)
    return _F32 if np.can_cast(dt, _F32, "safe") else _F64


_GUFUNC_IGNORED_KWARGS = ("order", "subok")  # accepted by NumPy gufuncs, no effect on values


class _GufuncKwargs:
    """The keyword arguments the reference forwards to its gufunc (`**kwargs` at
    numbagg/decorators.py:311-341, 380-414, 471-487, 712-731, 783-807, 836-876, 1069-1089):
    `out=`, `dtype=`, `casting=` (plus `order=` / `subok=`, which do not change values).
    Anything else raises the TypeError NumPy raises."""

    def __init__(self, name: str, kwargs: dict):
        self.name = name
        out = kwargs.pop("out", None)
        if isinstance(out, tuple):
            if len(out) != 1:
                raise ValueError("The 'out' tuple must have exactly one entry per ufunc output")
            out = out[0]
        self.out = out
        self.dtype = kwargs.pop("dtype", None)
        if self.dtype is not None:
            self.dtype = np.dtype(self.dtype)
        self.casting = kwargs.pop("casting", "same_kind")
        if self.casting not in ("no", "equiv", "safe", "same_kind", "unsafe"):
            raise ValueError("casting must be one of 'no', 'equiv', 'safe', 'same_kind', or 'unsafe'")
        for k in _GUFUNC_IGNORED_KWARGS:
            kwargs.pop(k, None)
        if kwargs:
            raise TypeError(f"{name}() got an unexpected keyword argument '{sorted(kwargs)[0]}'")

    def loop_dtype(self, natural: np.dtype, loops=(_F32, _F64), dtype_ok=None) -> np.dtype:
        """`natural`: the loop NumPy would pick from the inputs.  `dtype=` selects a loop by its
        output dtype; `dtype_ok` restricts which (the move / fill gufuncs also have int64 scalar
        operands, so only their float64 loop is reachable through `dtype=`)."""
        if self.dtype is None:
            return natural
        ok = loops if dtype_ok is None else dtype_ok
        if self.dtype not in ok:
            raise TypeError(f"No loop matching the specified signature and casting was found for ufunc {self.name}")
        return self.dtype

    def check_inputs(self, loop: np.dtype, *in_dtypes: np.dtype) -> None:
        for i, dt in enumerate(in_dtypes):
            if not np.can_cast(dt, loop, self.casting):
                raise TypeError(f"Cannot cast ufunc '{self.name}' input {i} from {dt!r} to {loop!r} "
                                f"with casting rule '{self.casting}'")

    def check_out(self, loop: np.dtype, shape: tuple) -> None:
        if self.out is None:
            return
        odt = dev.np_dtype_of(self.out)
        if not np.can_cast(loop, odt, self.casting):
            raise TypeError(f"Cannot cast ufunc '{self.name}' output from {loop!r} to {odt!r} "
                            f"with casting rule '{self.casting}'")
        if tuple(self.out.shape) != tuple(shape):
            raise ValueError(f"operands could not be broadcast together: output of shape {tuple(shape)} "
                             f"does not fit `out` of shape {tuple(self.out.shape)}")


def _on_tensor_device(fn):
    """Run a device-level entry on the device its tensors live on (not the caller's current
    device): pointers, workspaces and the stream handed to the C ABI must belong to one device.
    Tensors on different devices are an error."""
    import functools

    def _tensors(x, acc):
        if isinstance(x, torch.Tensor):
            if x.is_cuda:
                acc.append(x)
        elif isinstance(x, (list, tuple)):
            for y in x:
                _tensors(y, acc)

    @functools.wraps(fn)
    def wrapper(*args, **kwargs):
        ts: list[torch.Tensor] = []
        _tensors(args, ts)
        _tensors(tuple(kwargs.values()), ts)
        if not ts:
            return fn(*args, **kwargs)
        d = ts[0].device
        for t in ts[1:]:
            if t.device != d:
                raise ValueError(f"{fn.__name__}: operands live on different devices ({d} and {t.device})")
        if torch.cuda.current_device() == d.index:
            return fn(*args, **kwargs)
        with torch.cuda.device(d):
            return fn(*args, **kwargs)

    return wrapper


def _is_int(x) -> bool:
    return isinstance(x, (int, np.integer)) and not isinstance(x, (bool, np.bool_))


def _wants_tensor(*xs) -> bool:
    return any(dev.is_tensor(x) for x in xs)


def _finish(result: torch.Tensor, as_tensor: bool, out=None):
    """Return a tensor to tensor callers and a numpy array to numpy callers."""
    if out is not None:
        if dev.is_tensor(out):
            out.copy_(result)
            return out
        return dev.to_host(result, out)
    if as_tensor:
        return result
    return dev.to_host(result)


# ------------------------------------------------------- host <-> device pipelining (numpy)
_PIPELINE_MIN_BYTES = 64 << 20
_PIPELINE_CHUNK_BYTES = 96 << 20
_PIPELINE_STREAMS: dict[int, list] = {}


def _pipeline_streams(device: torch.device, k: int = 3):
    key = device.index if device.index is not None else torch.cuda.current_device()
    if key not in _PIPELINE_STREAMS:
        _PIPELINE_STREAMS[key] = [torch.cuda.Stream(device=device) for _ in range(k)]
    return _PIPELINE_STREAMS[key]


def _rows_pipelined(host_arrs: list[np.ndarray], work_dtype: np.dtype, axis: int, device_fn) -> np.ndarray | None:
    """numpy inputs whose leading dimension is NOT the core axis are independent along it:
    stream them through the GPU in row blocks on a few CUDA streams so that the H2D copy of
    block i+1, the kernels of block i and the D2H copy of block i-1 overlap (PCIe is full
    duplex).  `device_fn(list_of_device_tensors) -> device tensor` of the block's shape.
    Returns None when the input does not qualify (small, 1-D, core axis leading)."""
    a0 = host_arrs[0]
    nd = a0.ndim
    if nd < 2 or axis % nd == 0 or a0.shape[0] < 2:
        return None
    if any(x.shape != a0.shape or not x.flags.c_contiguous or x.dtype != work_dtype for x in host_arrs):
        return None
    if a0.nbytes < _PIPELINE_MIN_BYTES:
        return None
    device = dev.require_cuda()
    tdt = dev._NP_TO_TORCH[np.dtype(work_dtype)]
    nblocks = int(min(a0.shape[0], max(2, a0.nbytes // _PIPELINE_CHUNK_BYTES)))
    bounds = [a0.shape[0] * i // nblocks for i in range(nblocks + 1)]
    host_in = [torch.from_numpy(x) for x in host_arrs]
    host_out = torch.empty(a0.shape, dtype=tdt, pin_memory=True)
    streams = _pipeline_streams(device)
    main = torch.cuda.current_stream(device)
    for st in streams:
        st.wait_stream(main)
    for i in range(nblocks):
        st = streams[i % len(streams)]
        lo, hi = bounds[i], bounds[i + 1]
        with torch.cuda.stream(st):
            d_in = [h[lo:hi].to(device, non_blocking=True) for h in host_in]
            d_out = device_fn(d_in)
            host_out[lo:hi].copy_(d_out, non_blocking=True)
    for st in streams:
        main.wait_stream(st)
    main.synchronize()
    return host_out.numpy()


def _core_pipelined(host_arrs: list[np.ndarray], work_dtype: np.dtype, reverse: bool, chunk_fn) -> np.ndarray | None:
    """1-D numpy inputs (one long core axis): stream the axis through the GPU in chunks, in
    scan order, handing the running state from chunk to chunk ON THE DEVICE -- the same
    halo / carry protocol the multi-GPU shards use (include/nbg_b200.h) -- while the H2D copy
    of the next chunk and the D2H copy of the previous one overlap with the kernels.
    `chunk_fn(device_chunks, state) -> (device_out, new_state)`; `state` is None for the
    first chunk.  Returns None when the input does not qualify."""
    a0 = host_arrs[0]
    full_shape = a0.shape
    if a0.ndim == 0 or a0.size != a0.shape[-1] or a0.nbytes < _PIPELINE_MIN_BYTES:
        return None  # not a single long slice along the last axis
    if any(x.shape != a0.shape or not x.flags.c_contiguous or x.dtype != work_dtype for x in host_arrs):
        return None
    host_arrs = [x.reshape(-1) for x in host_arrs]
    a0 = host_arrs[0]
    device = dev.require_cuda()
    tdt = dev._NP_TO_TORCH[np.dtype(work_dtype)]
    n = a0.shape[0]
    nchunks = int(max(2, a0.nbytes // _PIPELINE_CHUNK_BYTES))
    bounds = [n * i // nchunks for i in range(nchunks + 1)]
    order = range(nchunks - 1, -1, -1) if reverse else range(nchunks)
    host_in = [torch.from_numpy(x) for x in host_arrs]
    host_out = torch.empty(a0.shape, dtype=tdt, pin_memory=True)
    main = torch.cuda.current_stream(device)
    copy_in, copy_out = _pipeline_streams(device)[:2]
    copy_in.wait_stream(main)
    copy_out.wait_stream(main)
    state = None
    staged = {}

    def stage(i):
        lo, hi = bounds[i], bounds[i + 1]
        with torch.cuda.stream(copy_in):
            t = [h[lo:hi].to(device, non_blocking=True) for h in host_in]
            ev = torch.cuda.Event()
            ev.record(copy_in)
        staged[i] = (t, ev)

    seq = list(order)
    stage(seq[0])
    for pos, i in enumerate(seq):
        if pos + 1 < len(seq):
            stage(seq[pos + 1])  # prefetch while chunk i computes
        d_in, ev = staged.pop(i)
        main.wait_event(ev)
        d_out, state = chunk_fn(d_in, state)
        for t in d_in:
            t.record_stream(main)
        done = torch.cuda.Event()
        done.record(main)
        copy_out.wait_event(done)
        with torch.cuda.stream(copy_out):
            host_out[bounds[i]:bounds[i + 1]].copy_(d_out, non_blocking=True)
        d_out.record_stream(copy_out)
    main.wait_stream(copy_out)
    main.synchronize()
    return host_out.numpy().reshape(full_shape)


class NumbaBase:
    """Counterpart of numbagg.decorators.NumbaBase (:103-162): carries the function's
    name/doc, ``repr`` and public signature.  ``target`` is always "cuda"."""

    def __init__(self, name: str, doc: str | None = None) -> None:
        self.__name__ = name
        self.__qualname__ = name
        self.__doc__ = doc
        self.__signature__ = inspect.signature(self.__call__)
        self.supports_parallel = True

    def __repr__(self) -> str:
        return f"numbagg.{self.__name__}"

    @property
    def target(self) -> str:
        return "cuda"

    def __call__(self, *args, **kwargs):
        raise NotImplementedError


# ------------------------------------------------------------------------------- moving
@_on_tensor_device
def run_move(name: str, arrs: list[torch.Tensor], window: int, min_count: int, axis: int,
             halos: list[torch.Tensor] | None = None) -> torch.Tensor:
    """Device-level entry (CUDA tensors of one float dtype, same shape) -> CUDA tensor."""
    view = dev.CoreView(arrs[0], axis)
    ts = [view.t] + [view.like(a) for a in arrs[1:]]
    out = torch.empty_like(view.t)
    halo_len = 0
    hs = [None, None]
    if halos:
        hv = [view.like(h) for h in halos]
        halo_len = hv[0].shape[view.axis]
        hs[: len(hv)] = hv
    rc = _lib.lib().nbg_move(
        _lib.MOVE_OPS[name], _NBG_DTYPE[dev.np_dtype_of(view.t)],
        dev.ptr(ts[0]), dev.ptr(ts[1]) if len(ts) > 1 else None, dev.ptr(out),
        view.outer, view.n, view.inner, int(window), int(min_count),
        dev.ptr(hs[0]), dev.ptr(hs[1]), halo_len, dev.stream_ptr(),
    )
    _lib.check(rc, f"nbg_move({name})")
    return view.restore(out)


class ndmove(NumbaBase):
    """N-dimensional moving-window function along one axis (numbagg ``ndmove``)."""

    def __init__(self, name: str, n_inputs: int = 1, doc: str | None = None):
        self.n_inputs = n_inputs
        super().__init__(name, doc)

    def __call__(self, *arr, window: int, min_count: int | None = None,
                 axis: int | tuple[int, ...] = -1, **kwargs):
        kw = _GufuncKwargs(self.__name__, kwargs)
        out = kw.out
        if len(arr) != self.n_inputs:
            raise TypeError(f"{self.__name__}() takes {self.n_inputs} array argument(s), got {len(arr)}")
        if min_count is None:
            min_count = window
        elif min_count < 0:
            raise ValueError(f"min_count must be positive: {min_count}")
        if isinstance(axis, tuple):
            if axis == ():
                if len(arr) > 1:
                    raise ValueError(
                        "`axis` cannot be an empty tuple when passing more than one array; since we default to returning the input."
                    )
                return arr[0]
            elif len(axis) > 1:
                raise ValueError(f"only one axis can be passed to {self.__name__}; got {axis}")
            (axis,) = axis
        as_tensor = _wants_tensor(*arr)
        arr = tuple(a if dev.is_tensor(a) else np.asarray(a) for a in arr)
        if not 0 < window <= arr[0].shape[axis]:
            raise ValueError(f"window not in valid range: {window}")
        if not _is_int(window) or not _is_int(min_count):
            # NumPy refuses to cast a float window/min_count to the int64 loop operand
            raise TypeError(f"window and min_count must be integers: {window!r}, {min_count!r}")
        in_dts = [dev.np_dtype_of(a) for a in arr]
        # `dtype=` also constrains the int64 window / min_count operands: only float64 resolves
        dt = kw.loop_dtype(_float_loop_dtype(*in_dts), dtype_ok=(_F64,))
        kw.check_inputs(dt, *in_dts)
        kw.check_out(dt, np.broadcast_shapes(*[tuple(a.shape) for a in arr]))
        if not as_tensor and out is None:
            piped = _rows_pipelined(list(arr), dt, axis,
                                    lambda ts: run_move(self.__name__, ts, window, min_count, axis))
            if piped is None:
                def chunk(ts, halos):
                    o = run_move(self.__name__, ts, window, min_count, 0, halos)
                    # next chunk's halo: the last `window` elements seen so far
                    new = [t[-window:] if t.shape[0] >= window else
                           (torch.cat([h, t])[-window:] if halos else t) for t, h in zip(ts, halos or ts)]
                    return o, [x.contiguous() for x in new]
                if axis % arr[0].ndim == arr[0].ndim - 1:
                    piped = _core_pipelined(list(arr), dt, False, chunk)
            if piped is not None:
                return piped
        ts = [dev.to_device(a, dt) for a in arr]
        if len(ts) > 1 and any(t.shape != ts[0].shape for t in ts):
            ts = [t.contiguous() for t in torch.broadcast_tensors(*ts)]
        res = run_move(self.__name__, ts, window, min_count, axis)
        return _finish(res, as_tensor, out)


# --------------------------------------------------------------------------- exp moving
@_on_tensor_device
def run_move_exp(name: str, arrs: list[torch.Tensor], alpha, min_weight: float, axis: int,
                 carry_in: torch.Tensor | None = None, want_agg: bool = False,
                 want_out: bool = True):
    """Device-level entry.  `alpha`: python float (scalar path) or a CUDA tensor that is 1-D
    of length n or has the data's shape.  Returns (out | None, agg | None)."""
    view = dev.CoreView(arrs[0], axis)
    ts = [view.t] + [view.like(a) for a in arrs[1:]]
    alpha_t, alpha_nd, alpha_scalar = None, 0, 0.0
    if dev.is_tensor(alpha):
        if alpha.dim() <= 1:
            alpha_t = alpha.contiguous()
        else:
            alpha_t, alpha_nd = view.like(alpha), 1
    else:
        alpha_scalar = float(alpha)
    L = _lib.lib()
    code = _lib.EXP_OPS[name]
    dcode = _NBG_DTYPE[dev.np_dtype_of(view.t)]
    if view.t.numel() == 0:  # empty input: empty result, identity aggregate (D = D2 = 1)
        agg0 = None
        if want_agg:
            agg0 = torch.zeros((view.outer * view.inner, _lib.NBG_EXP_STATE), dtype=torch.float64, device=view.t.device)
            agg0[:, :2] = 1.0
        return (view.restore(torch.empty_like(view.t)) if want_out else None), agg0
    out = torch.empty_like(view.t) if want_out else None
    slices = view.outer * view.inner
    agg = torch.empty((slices, _lib.NBG_EXP_STATE), dtype=torch.float64, device=view.t.device) if want_agg else None
    ws_bytes = L.nbg_move_exp_workspace_bytes(code, dcode, view.outer, view.n, view.inner)
    ws = torch.empty(max(ws_bytes, 1), dtype=torch.uint8, device=view.t.device)
    rc = L.nbg_move_exp(
        code, dcode, dev.ptr(ts[0]), dev.ptr(ts[1]) if len(ts) > 1 else None, dev.ptr(alpha_t),
        alpha_nd, alpha_scalar, float(min_weight), dev.ptr(out) if out is not None else None,
        view.outer, view.n, view.inner, dev.ptr(carry_in), dev.ptr(agg), ws.data_ptr(), ws_bytes,
        dev.stream_ptr(),
    )
    _lib.check(rc, f"nbg_move_exp({name})")
    return (view.restore(out) if out is not None else None), agg


class ndmoveexp(NumbaBase):
    """Exponentially-weighted moving function (numbagg ``ndmoveexp``).  ``alpha`` is a
    scalar, a 1-D array over the core axis, or an array with the data's shape."""

    def __init__(self, name: str, n_inputs: int = 1, doc: str | None = None):
        self.n_inputs = n_inputs
        super().__init__(name, doc)

    def __call__(self, *arr, alpha, min_weight: float = 0,
                 axis: int | tuple[int, ...] = -1, **kwargs):
        kw = _GufuncKwargs(self.__name__, kwargs)
        out = kw.out
        if len(arr) != self.n_inputs:
            raise TypeError(f"{self.__name__}() takes {self.n_inputs} array argument(s), got {len(arr)}")
        if isinstance(axis, tuple):
            if axis == ():
                if len(arr) > 1:
                    raise ValueError(
                        "`axis` cannot be an empty tuple when passing more than one array; since we default to returning the input."
                    )
                return arr[0]
            if len(axis) > 1:
                raise ValueError(f"Only one axis can be passed to {self.__name__}; got {axis}")
            (axis,) = axis
        as_tensor = _wants_tensor(*arr)
        arr = tuple(a if dev.is_tensor(a) else np.asarray(a) for a in arr)
        n = arr[0].shape[axis]
        alpha_is_array = isinstance(alpha, np.ndarray) or dev.is_tensor(alpha)
        if alpha_is_array:
            alpha_dtype = dev.np_dtype_of(alpha)
            if alpha.ndim == 0:
                alpha_is_array = False
                alpha = alpha.item() if alpha_dtype != _F32 else np.float32(alpha.item())
        if not alpha_is_array:
            # np.broadcast_to(alpha, n) in the reference: the scalar's own dtype takes part in
            # loop selection (python float -> float64, np.float32 -> float32)
            alpha_dtype = np.asarray(alpha).dtype
        in_dts = [dev.np_dtype_of(a) for a in arr]
        dt = kw.loop_dtype(_float_loop_dtype(*in_dts, alpha_dtype))
        kw.check_inputs(dt, *in_dts)
        kw.check_out(dt, np.broadcast_shapes(*[tuple(a.shape) for a in arr]))
        if not as_tensor and out is None and not alpha_is_array:
            a_s = float(np.float32(alpha)) if dt == _F32 else float(alpha)
            m_s = float(np.float32(min_weight)) if dt == _F32 else float(min_weight)
            piped = _rows_pipelined(list(arr), dt, axis,
                                    lambda ts: run_move_exp(self.__name__, ts, a_s, m_s, axis)[0])
            if piped is None and axis % arr[0].ndim == arr[0].ndim - 1:
                piped = _core_pipelined(list(arr), dt, False,
                                        lambda ts, st: run_move_exp(self.__name__, ts, a_s, m_s, 0, st, True, True))
            if piped is not None:
                return piped
        ts = [dev.to_device(a, dt) for a in arr]
        if len(ts) > 1 and any(t.shape != ts[0].shape for t in ts):
            ts = [t.contiguous() for t in torch.broadcast_tensors(*ts)]
        if alpha_is_array:
            if alpha.ndim == 1:
                if alpha.shape[0] != n:
                    raise ValueError(
                        f"alpha has length {alpha.shape[0]} but the core axis has length {n}"
                    )
                alpha_dev = dev.to_device(alpha, dt)
            else:
                alpha_dev = dev.to_device(alpha, dt)
                if alpha_dev.shape != ts[0].shape:
                    alpha_dev = alpha_dev.broadcast_to(ts[0].shape).contiguous()
        else:
            # the loop sees alpha rounded to the loop dtype
            alpha_dev = float(np.float32(alpha)) if dt == _F32 else float(alpha)
        mw = float(np.float32(min_weight)) if dt == _F32 else float(min_weight)
        res, _ = run_move_exp(self.__name__, ts, alpha_dev, mw, axis)
        return _finish(res, as_tensor, out)


# -------------------------------------------------------------------------------- fills
@_on_tensor_device
def run_fill(name: str, t: torch.Tensor, limit: int, axis: int,
             carry_in: torch.Tensor | None = None, want_agg: bool = False, want_out: bool = True):
    """Device-level entry for ffill/bfill on a float32/float64 CUDA tensor."""
    view = dev.CoreView(t, axis)
    L = _lib.lib()
    itemsize = view.t.element_size()
    if view.t.numel() == 0:  # nothing to scan: empty result, identity aggregate
        agg0 = torch.zeros((view.outer * view.inner, _lib.NBG_FILL_STATE), dtype=torch.int64, device=t.device) if want_agg else None
        return (view.restore(torch.empty_like(view.t)) if want_out else None), agg0
    out = torch.empty_like(view.t) if want_out else None
    slices = view.outer * view.inner
    agg = torch.empty((slices, _lib.NBG_FILL_STATE), dtype=torch.int64, device=t.device) if want_agg else None
    ws_bytes = L.nbg_fill_workspace_bytes(itemsize, view.outer, view.n, view.inner)
    ws = torch.empty(max(ws_bytes, 1), dtype=torch.uint8, device=t.device)
    rc = L.nbg_fill(
        _lib.FILL_DIRS[name], itemsize, dev.ptr(view.t), dev.ptr(out) if out is not None else None,
        view.outer, view.n, view.inner, int(limit), dev.ptr(carry_in), dev.ptr(agg),
        ws.data_ptr(), ws_bytes, dev.stream_ptr(),
    )
    _lib.check(rc, f"nbg_fill({name})")
    return (view.restore(out) if out is not None else None), agg


def fill_sentinel_bits(itemsize: int) -> int:
    """NaN payload that marks "depends on the predecessors' carry" in a sharded fill."""
    return int(_lib.lib().nbg_fill_sentinel_bits(int(itemsize)))


@_on_tensor_device
def run_fill_patch(name: str, out: torch.Tensor, limit: int, axis: int, carry: torch.Tensor) -> None:
    """Second step of the single-pass sharded fill: rewrite the sentinel run of `out` in place from
    the folded carry ((slices, NBG_FILL_STATE) int64).  The core axis must be the last one."""
    nd = out.dim()
    if axis % nd != nd - 1 or not out.is_contiguous():
        raise ValueError("run_fill_patch needs a C-contiguous tensor with the core axis last")
    n = out.shape[-1]
    outer = out.numel() // max(n, 1)
    rc = _lib.lib().nbg_fill_patch(_lib.FILL_DIRS[name], out.element_size(), dev.ptr(out), outer, n, 1, int(limit),
                                   dev.ptr(carry), dev.stream_ptr())
    _lib.check(rc, f"nbg_fill_patch({name})")


class ndfill(NumbaBase):
    """Forward/backward fill along one axis (numbagg ``ndfill``)."""

    def __call__(self, arr, *, limit: None | int = None, axis: int = -1, **kwargs):
        kw = _GufuncKwargs(self.__name__, kwargs)
        out = kw.out
        as_tensor = dev.is_tensor(arr)
        if not as_tensor:
            arr = np.asarray(arr)
        if limit is None:
            limit = arr.shape[axis]
        if limit < 0:
            raise ValueError(f"`limit` must be positive: {limit}")
        dt = dev.np_dtype_of(arr)
        if not np.issubdtype(dt, np.number):
            raise TypeError(f"Unsupported dtype for fill operation: {dt}")
        # the fill gufunc is compiled per dtype (decorators.py:427-463): the only loop is the input's own
        kw.loop_dtype(dt, dtype_ok=(dt,) if dt == _F64 else ())
        kw.check_out(dt, tuple(arr.shape))
        if dt.kind in "iu":
            # np.isnan is constant-false for integers (funcs.py:303,319): the loop is a copy
            res = arr.clone() if as_tensor else arr.copy()
            if out is not None:
                out[...] = res
                return out
            return res
        if dt.kind == "c":
            raise TypeError(f"Unsupported dtype for fill operation: {dt}")
        work = _F32 if dt == np.dtype(np.float16) else dt  # float16 <-> float32 is exact
        if not as_tensor and out is None and work == dt:
            piped = _rows_pipelined([arr], dt, axis, lambda ts: run_fill(self.__name__, ts[0], limit, axis)[0])
            if piped is None and axis % arr.ndim == arr.ndim - 1:
                piped = _core_pipelined([arr], dt, self.__name__ == "bfill",
                                        lambda ts, st: run_fill(self.__name__, ts[0], limit, 0, st, True, True))
            if piped is not None:
                return piped
        t = dev.to_device(arr, work)
        if t.dim() == 0:
            raise ValueError("ffill/bfill need at least one dimension")
        res, _ = run_fill(self.__name__, t, limit, axis)
        if work != dt:
            res = res.to(dev._NP_TO_TORCH[dt])
        return _finish(res, as_tensor, out)


# ------------------------------------------------------------------------------ grouped
@_on_tensor_device
def run_group(name: str, values: torch.Tensor, labels: torch.Tensor, num_labels: int, ddof: int,
              labels_per_row: bool = False) -> torch.Tensor:
    """Device-level entry: values (rows, n) contiguous, labels (n,) or (rows, n), both CUDA;
    values dtype in {f32, f64, i32, i64}, labels in {i32, i64} -> (rows, num_labels)."""
    L = _lib.lib()
    code = _lib.GROUP_OPS[name]
    vcode = _NBG_DTYPE[dev.np_dtype_of(values)]
    lcode = _NBG_DTYPE[dev.np_dtype_of(labels)]
    rows, n = values.shape
    out = torch.empty((rows, num_labels), dtype=values.dtype, device=values.device)
    ws_bytes = L.nbg_group_workspace_bytes(code, vcode, rows, n, num_labels)
    ws = torch.empty(max(ws_bytes, 1), dtype=torch.uint8, device=values.device)
    rc = L.nbg_group(
        code, vcode, lcode, dev.ptr(values), dev.ptr(labels), int(labels_per_row), dev.ptr(out),
        rows, n, int(num_labels), int(ddof), ws.data_ptr(), ws_bytes, dev.stream_ptr(),
    )
    _lib.check(rc, f"nbg_group({name})")
    return out


class groupndreduce(NumbaBase):
    """N-dimensional grouped aggregation (numbagg ``groupndreduce``): ``axis=int`` (labels
    1-D along that axis), ``axis=tuple`` (labels shaped like those axes) or ``axis=None``
    (labels shaped like values).  The group axis is the LAST axis of the result."""

    def __init__(self, name: str, *, supports_ddof: bool = False, supports_bool: bool = True,
                 supports_ints: bool = True, doc: str | None = None) -> None:
        self.supports_bool = supports_bool
        self.supports_ints = supports_ints
        self.supports_ddof = supports_ddof
        super().__init__(name, doc)

    def __call__(self, values, labels, *, ddof: int = 1, num_labels: int | None = None,
                 axis: int | tuple[int, ...] | None = None):
        as_tensor = _wants_tensor(values, labels)
        if not dev.is_tensor(values):
            values = np.asarray(values)
        if not dev.is_tensor(labels):
            labels = np.asarray(labels)
        vdt = dev.np_dtype_of(values)
        ldt = dev.np_dtype_of(labels)
        if ldt.kind not in "i":
            raise TypeError(
                "labels must be an integer array; it's expected to have already been factorized with a function such as `pd.factorize`"
            )
        vsize = math.prod(values.shape)
        # The reference widens labels so that a per-group count cannot overflow
        # (decorators.py:581-596).  Counts here are always 64-bit; labels only need a dtype
        # the kernels read (int32/int64).
        if vdt == np.bool_:
            if not self.supports_bool:
                raise TypeError(
                    f"{self.__name__} does not support boolean input. Convert to a numeric type first."
                )
            vdt = np.dtype(np.int32)
        if num_labels is None:
            # `int` so that a label at the dtype's maximum doesn't overflow the add.
            num_labels = int(labels.max()) + 1
        if not self.supports_ints and np.issubdtype(vdt, np.integer):
            work_dt = result_dt = _F64
        else:
            result_dt = vdt
            if vdt in _NBG_DTYPE:
                work_dt = vdt
            elif vdt.kind in "iu":
                work_dt = np.dtype(np.int64)  # narrow ints: wrap back to result_dt at the end
            elif vdt == np.dtype(np.float16):
                work_dt = result_dt = _F32
            else:
                raise TypeError(f"unsupported values dtype {vdt}")

        vshape = tuple(values.shape)
        lshape = tuple(labels.shape)
        nd = len(vshape)
        if axis is None:
            if vshape != lshape:
                raise ValueError(
                    "axis required if values and labels have different "
                    f"shapes: {vshape} vs {lshape}"
                )
            core_axes = tuple(range(nd))
        elif isinstance(axis, (int, np.integer)):
            if lshape != (vshape[axis],):
                raise ValueError(
                    "values must have same shape along axis as labels: "
                    f"{(vshape[axis],)} vs {lshape}"
                )
            core_axes = (int(axis) % nd,)
        else:
            values_shape = tuple(vshape[ax] for ax in axis)
            if lshape != values_shape:
                raise ValueError(
                    "values must have same shape along axis as labels: "
                    f"{values_shape} vs {lshape}"
                )
            core_axes = tuple(int(ax) % nd for ax in axis)

        v = dev.to_device(values, work_dt)
        lab = dev.to_device(labels, np.dtype(np.int64) if ldt.itemsize > 4 else np.dtype(np.int32))
        if len(core_axes) and core_axes != tuple(range(nd - len(core_axes), nd)):
            v = torch.movedim(v, core_axes, tuple(range(nd - len(core_axes), nd)))
        bshape = tuple(v.shape[: nd - len(core_axes)])
        n = math.prod(v.shape[nd - len(core_axes):])
        v2 = v.reshape(-1, n) if v.is_contiguous() else v.contiguous().reshape(-1, n)
        res = run_group(self.__name__, v2, lab.reshape(-1).contiguous(), num_labels,
                        ddof if self.supports_ddof else 1)
        res = res.reshape(bshape + (num_labels,))
        if work_dt != result_dt:
            res = res.to(dev._NP_TO_TORCH[result_dt])
        return _finish(res, as_tensor)


# ------------------------------------------------------- grouped: three-step (sharded) form
def group_record_words(name: str) -> int:
    """8-byte slots per (row, label) record of this op's accumulator state (1, 2 or 4)."""
    return int(_lib.lib().nbg_group_record_words(_lib.GROUP_OPS[name]))


@_on_tensor_device
def run_group_partial(name: str, values: torch.Tensor, labels: torch.Tensor, num_labels: int,
                      index_offset: int = 0, labels_per_row: bool = False) -> torch.Tensor:
    """init + accumulate for one element shard.  Returns the accumulator state as an int64
    tensor (rows, num_labels, record_words) -- raw 8-byte slots (include/nbg_b200.h)."""
    L = _lib.lib()
    code = _lib.GROUP_OPS[name]
    vcode = _NBG_DTYPE[dev.np_dtype_of(values)]
    lcode = _NBG_DTYPE[dev.np_dtype_of(labels)]
    rows, n = values.shape
    ws_bytes = L.nbg_group_workspace_bytes(code, vcode, rows, n, num_labels)
    ws = torch.empty(max(ws_bytes, 8), dtype=torch.uint8, device=values.device)
    assert ws.data_ptr() % 256 == 0
    _lib.check(L.nbg_group_init(code, vcode, ws.data_ptr(), rows, num_labels, dev.stream_ptr()), "nbg_group_init")
    _lib.check(
        L.nbg_group_accumulate(code, vcode, lcode, dev.ptr(values), dev.ptr(labels), int(labels_per_row),
                               ws.data_ptr(), ws_bytes, rows, n, num_labels, int(index_offset), dev.stream_ptr()),
        f"nbg_group_accumulate({name})",
    )
    words = group_record_words(name)
    return ws[: words * rows * num_labels * 8].view(torch.int64).view(rows, num_labels, words)


def group_state_channels(name: str, state: torch.Tensor, rows: int, num_labels: int) -> list[torch.Tensor]:
    """Writable 1-D views (rows * num_labels int64 words each) of channels 0..2 of a group state
    returned by run_group_partial -- what the multi-GPU combine all-reduces.  Channel meanings per
    op: DESIGN.md "group workspace"; channels an op does not use alias another one."""
    import ctypes

    lay = (ctypes.c_int64 * 5)()
    _lib.check(_lib.lib().nbg_group_record_layout(_lib.GROUP_OPS[name], rows, num_labels, lay), "nbg_group_record_layout")
    flat = state.reshape(-1)
    slots = rows * num_labels
    return [torch.as_strided(flat, (slots,), (int(lay[0]),), int(lay[1 + c])) for c in range(3)]


@_on_tensor_device
def run_group_combine(name: str, vdtype: np.dtype, acc: torch.Tensor, other: torch.Tensor) -> None:
    """acc <- merge(acc, other) where `other` covers LATER elements (nbg_group_combine)."""
    rows, K, _ = acc.shape
    rc = _lib.lib().nbg_group_combine(_lib.GROUP_OPS[name], _NBG_DTYPE[np.dtype(vdtype)], acc.data_ptr(),
                                      other.data_ptr(), rows, K, dev.stream_ptr())
    _lib.check(rc, f"nbg_group_combine({name})")


@_on_tensor_device
def run_group_finalize(name: str, vdtype: np.dtype, state: torch.Tensor, ddof: int) -> torch.Tensor:
    rows, K, _ = state.shape
    out = torch.empty((rows, K), dtype=dev._NP_TO_TORCH[np.dtype(vdtype)], device=state.device)
    rc = _lib.lib().nbg_group_finalize(_lib.GROUP_OPS[name], _NBG_DTYPE[np.dtype(vdtype)], state.data_ptr(),
                                       out.data_ptr(), rows, K, int(ddof), dev.stream_ptr())
    _lib.check(rc, f"nbg_group_finalize({name})")
    return out


# ------------------------------------------------------------------- plain reductions
_REDUCE_FLOAT_ONLY = ("nanmean", "nanvar", "nanstd")
_REDUCE_EMPTY_ERRORS = {
    "nanargmax": "All-NaN slice encountered",
    "nanargmin": "All-NaN slice encountered",
    "nanmax": "zero-size array to reduction operation fmax which has no identity",
    "nanmin": "zero-size array to reduction operation fmin which has no identity",
}


def _reduce_loop_dtype(name: str, dt: np.dtype) -> np.dtype:
    """The gufunc loop NumPy would pick among the reference's signatures (funcs.py:23-242):
    int32 / int64 / float32 / float64, or float32 / float64 only for nanmean/nanvar/nanstd."""
    if dt.kind == "f":
        return _F32 if dt.itemsize <= 4 else _F64
    if name in _REDUCE_FLOAT_ONLY:
        # first loop the input casts to safely: 8/16-bit integers reach float32.  bool does so for
        # nanmean only -- with the integer `ddof` operand of nanvar / nanstd NumPy resolves bool to
        # the float64 loop (probed against the reference)
        small = dt.kind in "iu" and dt.itemsize <= 2
        if small or (dt.kind == "b" and name == "nanmean"):
            return _F32
        return _F64
    if dt.kind == "b" or (dt.kind in "iu" and dt.itemsize < 4) or dt == np.dtype(np.int32):
        return np.dtype(np.int32)
    if dt.kind in "iu" and dt != np.dtype(np.uint64):
        return np.dtype(np.int64)
    if dt == np.dtype(np.uint64):
        return _F64
    raise TypeError(f"Unsupported dtype for {name}: {dt}")


def _reduce_out_dtype(name: str, work: np.dtype) -> torch.dtype:
    if name in ("allnan", "anynan"):
        return torch.uint8
    if name in ("nancount", "nanargmax", "nanargmin"):
        return torch.int64
    if name in ("nanmax", "nanmin") and work.kind == "i":
        return torch.int64
    return dev._NP_TO_TORCH[work]


class ReduceView:
    """(outer, n, inner) C-contiguous view of `t` whose middle axis is the flattened block of
    reduced axes, taken in the order given (that order defines the flat index nanarg* return
    and the reference's accumulation order).  No copy when the reduced axes are adjacent in
    memory in that order -- axis=None, one axis, or a run of neighbouring axes of a C-, F- or
    otherwise permuted-contiguous array; anything else is permuted with one copy."""

    def __init__(self, t: torch.Tensor, axes: tuple[int, ...]):
        nd = t.dim()
        # memory order of the dims, outermost first; size-1 dims sort anywhere
        perm = sorted(range(nd), key=lambda d: (-t.stride(d) if t.shape[d] > 1 else 0, d))
        if nd and not t.permute(perm).is_contiguous():
            t = t.contiguous()
            perm = list(range(nd))
        pos = {d: i for i, d in enumerate(perm)}
        red = [pos[ax] for ax in axes]
        batch_dims = [d for d in range(nd) if d not in axes]  # output dims, original order
        shape_p = [t.shape[d] for d in perm]
        adjacent = all(red[i + 1] == red[i] + 1 for i in range(len(red) - 1))
        # size-1 dims can be anywhere in `perm`; the batch dims must keep a consistent order
        batch_p = [d for d in perm if d not in axes]
        if red and adjacent:
            self.t = t.permute(perm)
            self.outer = math.prod(shape_p[: red[0]])
            self.n = math.prod(shape_p[red[0] : red[-1] + 1])
            self.inner = math.prod(shape_p[red[-1] + 1 :])
        else:
            self.t = t.permute(batch_p + list(axes)).contiguous()
            self.outer = math.prod(t.shape[d] for d in batch_p)
            self.n = math.prod(t.shape[ax] for ax in axes)
            self.inner = 1
        self._batch_shape_p = [t.shape[d] for d in batch_p]
        # permutation taking the (memory-ordered) batch dims back to their original order
        self._back = [batch_p.index(d) for d in batch_dims]

    def restore(self, flat: torch.Tensor) -> torch.Tensor:
        out = flat.reshape(self._batch_shape_p)
        return out.permute(self._back) if self._back else out


@_on_tensor_device
def run_reduce(name: str, t: torch.Tensor, axes: tuple[int, ...], ddof: int = 1) -> torch.Tensor:
    """Device-level entry: reduce `axes` (in that order) of a CUDA tensor whose dtype is one
    of the loop dtypes; returns a CUDA tensor of the batch shape."""
    work = dev._TORCH_TO_NP[t.dtype]
    view = ReduceView(t, axes)
    L = _lib.lib()
    op = _lib.REDUCE_OPS[name]
    outs = view.outer * view.inner
    out = torch.empty(outs, dtype=_reduce_out_dtype(name, work), device=t.device)
    ws_bytes = L.nbg_reduce_workspace_bytes(op, _NBG_DTYPE[work], view.outer, view.n, view.inner)
    ws = torch.empty(max(ws_bytes, 1), dtype=torch.uint8, device=t.device)
    rc = L.nbg_reduce(op, _NBG_DTYPE[work], dev.ptr(view.t), dev.ptr(out), view.outer, view.n, view.inner,
                      int(ddof), ws.data_ptr(), ws_bytes, dev.stream_ptr())
    _lib.check(rc, f"nbg_reduce({name})")
    if name in ("allnan", "anynan"):
        out = out.view(torch.bool)
    return view.restore(out)


@_on_tensor_device
def run_reduce_partial(name: str, t: torch.Tensor, axes: tuple[int, ...], index_offset: int = 0):
    """Element shard -> (3, outs) int64 state records (include/nbg_b200.h) + the view."""
    work = dev._TORCH_TO_NP[t.dtype]
    view = ReduceView(t, axes)
    L = _lib.lib()
    op = _lib.REDUCE_OPS[name]
    outs = view.outer * view.inner
    states = torch.empty((_lib.NBG_REDUCE_STATE_WORDS, outs), dtype=torch.int64, device=t.device)
    ws_bytes = L.nbg_reduce_workspace_bytes(op, _NBG_DTYPE[work], view.outer, view.n, view.inner)
    ws = torch.empty(max(ws_bytes, 1), dtype=torch.uint8, device=t.device)
    rc = L.nbg_reduce_partial(op, _NBG_DTYPE[work], dev.ptr(view.t), dev.ptr(states), view.outer, view.n,
                              view.inner, int(index_offset), ws.data_ptr(), ws_bytes, dev.stream_ptr())
    _lib.check(rc, f"nbg_reduce_partial({name})")
    return states, view


@_on_tensor_device
def run_reduce_merge(name: str, work: np.dtype, states: torch.Tensor, n_total: int, ddof: int = 1) -> torch.Tensor:
    """Fold (parts, 3, outs) gathered state records and finalize -> flat (outs,) result."""
    parts, _, outs = states.shape
    L = _lib.lib()
    out = torch.empty(outs, dtype=_reduce_out_dtype(name, work), device=states.device)
    rc = L.nbg_reduce_merge(_lib.REDUCE_OPS[name], _NBG_DTYPE[work], dev.ptr(states), parts, outs, dev.ptr(out),
                            int(n_total), int(ddof), dev.stream_ptr())
    _lib.check(rc, f"nbg_reduce_merge({name})")
    return out.view(torch.bool) if name in ("allnan", "anynan") else out


def _normalize_axes(axis, nd: int) -> tuple[int, ...]:
    if axis is None:
        return tuple(range(nd))
    if not isinstance(axis, tuple):
        axis = (axis,)
    return tuple(np.lib.array_utils.normalize_axis_tuple(axis, nd))


def _reduce_result(res: torch.Tensor, as_tensor: bool):
    if as_tensor:
        return res
    host = dev.to_host(res)
    return host[()] if host.ndim == 0 else host


class ndaggregate(NumbaBase):
    """Simple aggregations over one or more axes (numbagg ``ndaggregate``,
    decorators.py:188-260): allnan, anynan, nancount, nansum, nanmean, nanvar, nanstd."""

    def __init__(self, name: str, supports_ddof: bool = False, doc: str | None = None):
        self.supports_ddof = supports_ddof
        super().__init__(name, doc)
        if not supports_ddof:
            params = [p for p in self.__signature__.parameters.values() if p.name != "ddof"]
            self.__signature__ = self.__signature__.replace(parameters=params)

    def __call__(self, *arrays, ddof: int = 1, axis: int | tuple[int, ...] | None = None):
        if not all(isinstance(a, np.ndarray) or dev.is_tensor(a) for a in arrays):
            raise TypeError(f"All positional arguments to {self} must be arrays: {arrays}")
        if len(arrays) != 1:
            raise TypeError(f"{self.__name__}() takes exactly one array ({len(arrays)} given)")
        (arr,) = arrays
        as_tensor = dev.is_tensor(arr)
        nd = arr.dim() if as_tensor else arr.ndim
        axes = _normalize_axes(axis, nd)
        dt = dev.np_dtype_of(arr)
        work = _reduce_loop_dtype(self.__name__, dt)
        t = dev.to_device(arr, work)
        if len(axes) > 1:
            # decorators.py:211-230: reduce in memory order, largest stride first
            order = np.argsort([t.stride(ax) for ax in axes], kind="stable")[::-1]
            axes = tuple(axes[i] for i in order)
        res = run_reduce(self.__name__, t, axes, ddof if self.supports_ddof else 1)
        return _reduce_result(res, as_tensor)


class ndreduce(NumbaBase):
    """Reductions written as scalar-returning loops (numbagg ``ndreduce``,
    decorators.py:906-1031): nanargmax, nanargmin, nanmax, nanmin.  A tuple `axis` is reduced
    in the order given; nanarg* return the flat index within those axes."""

    def __call__(self, arr, *args, axis: tuple[int, ...] | int | None = None):
        if args:
            raise TypeError(f"{self.__name__}() takes one positional argument")
        as_tensor = dev.is_tensor(arr)
        if not as_tensor:
            arr = np.asarray(arr)
        nd = arr.dim() if as_tensor else arr.ndim
        axes = _normalize_axes(axis, nd)
        name = self.__name__
        shape = tuple(arr.shape)
        n = math.prod(shape[ax] for ax in axes)
        rows = math.prod(s for d, s in enumerate(shape) if d not in axes)
        if n == 0 and rows > 0:
            raise ValueError(_REDUCE_EMPTY_ERRORS[name])
        work = _reduce_loop_dtype(name, dev.np_dtype_of(arr))
        t = dev.to_device(arr, work)
        res = run_reduce(name, t, axes)
        if name in ("nanargmax", "nanargmin"):
            if as_tensor:
                if bool((res < 0).any()):
                    raise ValueError("All-NaN slice encountered")
                return res
            host = dev.to_host(res)
            if (host < 0).any():
                raise ValueError("All-NaN slice encountered")
            return host[()] if host.ndim == 0 else host
        return _reduce_result(res, as_tensor)


# ------------------------------------------------------------------------- quantiles
@_on_tensor_device
def run_quantile(t: torch.Tensor, q: torch.Tensor, axes: tuple[int, ...]) -> torch.Tensor:
    """Device-level entry: float64 CUDA tensor `t`, float64 CUDA vector `q` (quantiles in
    [0, 1] or NaN); returns (len(q),) + batch shape."""
    view = ReduceView(t, axes)
    cube = view.t.reshape(view.outer, view.n, view.inner)
    if view.inner != 1:  # selection needs each slice contiguous: one transposing copy
        cube = cube.permute(0, 2, 1).contiguous()
    rows = view.outer * view.inner
    L = _lib.lib()
    m_all = int(q.numel())
    flat = cube.reshape(rows, view.n)
    out = torch.empty((rows, m_all), dtype=torch.float64, device=t.device)
    # the C entry takes <= 16 quantiles and (on its long-row path) <= 65535 rows per call; the long-row
    # workspace (histograms of 2048 words per target + a 2048-key candidate list per row) is kept at
    # <= ~256 MB by shrinking the row block with the number of quantiles, and it is allocated once
    m_max = min(m_all, _lib.NBG_QUANTILE_MAX_Q)
    row_step = rows
    if view.n > 4096 and rows and m_all:
        per_row = max(1, int(L.nbg_quantile_workspace_bytes(1024, view.n, max(m_max, 1))) // 1024)
        row_step = max(64, min(32768, (256 << 20) // per_row))
    ws_cap = int(L.nbg_quantile_workspace_bytes(min(rows, row_step), view.n, max(m_max, 1))) if rows and m_all else 0
    ws_all = torch.empty(max(ws_cap, 1), dtype=torch.uint8, device=t.device)
    for r0 in range(0, rows, max(row_step, 1)):
        block = flat[r0:r0 + row_step]
        nrows = int(block.shape[0])
        for lo in range(0, m_all, _lib.NBG_QUANTILE_MAX_Q):
            qc = q[lo:lo + _lib.NBG_QUANTILE_MAX_Q].contiguous()
            m = int(qc.numel())
            part = torch.empty((nrows, m), dtype=torch.float64, device=t.device)
            ws_bytes = L.nbg_quantile_workspace_bytes(nrows, view.n, m)
            ws = ws_all if ws_bytes <= ws_all.numel() else torch.empty(ws_bytes, dtype=torch.uint8, device=t.device)
            rc = L.nbg_quantile(dev.ptr(block), dev.ptr(qc), dev.ptr(part), nrows, view.n, m, ws.data_ptr(),
                                ws_bytes, dev.stream_ptr())
            _lib.check(rc, "nbg_quantile")
            out[r0:r0 + nrows, lo:lo + m] = part
    batch_shape = tuple(view.restore(out.new_empty(rows)).shape)
    if m_all == 0:
        return out.new_empty((0,) + batch_shape)
    return torch.stack([view.restore(out[:, i].contiguous()) for i in range(m_all)])


class ndquantile(NumbaBase):
    """numbagg ``ndquantile`` (decorators.py:821-884): ``nanquantile(a, quantiles, axis=None)``;
    the quantile axis comes first in the result, a scalar `quantiles` is squeezed away."""

    def __call__(self, a, quantiles, axis: int | tuple[int, ...] | None = None, **kwargs):
        from collections.abc import Iterable

        kw = _GufuncKwargs(self.__name__, kwargs)
        squeeze = not isinstance(quantiles, Iterable)
        qs = np.asarray([quantiles] if squeeze else quantiles, dtype=np.float64)
        if qs.ndim != 1:
            raise ValueError("quantiles must be a scalar or one-dimensional")
        if any(qs < 0) or any(qs > 1):
            raise ValueError(f"quantiles must be in the range [0, 1], inclusive. Got {qs}.")
        as_tensor = dev.is_tensor(a)
        if not as_tensor:
            a = np.asarray(a)
        nd = a.dim() if as_tensor else a.ndim
        axes = _normalize_axes(axis, nd)
        dt = dev.np_dtype_of(a)
        if dt.kind not in "fiub":
            raise TypeError(f"Unsupported dtype for {self.__name__}: {dt}")
        kw.loop_dtype(_F64, dtype_ok=(_F64,))
        kw.check_inputs(_F64, dt)
        t = dev.to_device(a, _F64)  # the reference has a float64 loop only
        q = torch.from_numpy(qs).to(t.device)
        res = run_quantile(t, q, axes)
        if kw.out is not None:
            # the gufunc's output is (..., m) BEFORE the quantile axis is moved first and squeezed
            # (decorators.py:876-884); `out` has that layout
            raw = torch.movedim(res, 0, -1)
            kw.check_out(_F64, tuple(raw.shape))
            _finish(raw, as_tensor, kw.out)
        if squeeze:
            res = res[0]
        return _reduce_result(res, as_tensor)


# ------------------------------------------------------------------- matrix functions
@_on_tensor_device
def run_matrix(name: str, t: torch.Tensor, *, window: int = 0, min_count: int = 0, alpha: torch.Tensor | None = None,
               min_weight: float = 0.0) -> torch.Tensor:
    """Device-level entry for the six matrix functions.  Static ops take (..., vars, obs),
    the moving / exponential ones (..., obs, vars); `alpha` is (obs,) or batch-shaped + (obs,)
    in the dtype of `t`."""
    static = name in ("nancorrmatrix", "nancovmatrix")
    t = t if t.is_contiguous() else t.contiguous()
    batch_shape = tuple(t.shape[:-2])
    batch = math.prod(batch_shape)
    if static:
        nv, no = t.shape[-2:]
        out = torch.empty(batch_shape + (nv, nv), dtype=t.dtype, device=t.device)
    else:
        no, nv = t.shape[-2:]
        out = torch.empty(batch_shape + (no, nv, nv), dtype=t.dtype, device=t.device)
    per_item = 0
    if alpha is not None:
        if alpha.dim() > 1:
            alpha = alpha.expand(batch_shape + (no,))
            per_item = 1
        alpha = alpha.contiguous()
    rc = _lib.lib().nbg_matrix(_lib.MATRIX_OPS[name], _NBG_DTYPE[dev._TORCH_TO_NP[t.dtype]], dev.ptr(t), dev.ptr(alpha),
                               per_item, float(min_weight), dev.ptr(out), batch, no, nv, int(window), int(min_count),
                               dev.stream_ptr())
    _lib.check(rc, f"nbg_matrix({name})")
    return out


def _matrix_loop_dtype(*dtypes: np.dtype) -> np.dtype:
    return _float_loop_dtype(*dtypes)


class ndmatrix(NumbaBase):
    """numbagg ``ndmatrix`` (decorators.py:677-740): ``(..., vars, obs) -> (..., vars, vars)``."""

    def __call__(self, a, **kwargs):
        kw = _GufuncKwargs(self.__name__, kwargs)
        as_tensor = dev.is_tensor(a)
        nd = a.dim() if as_tensor else np.ndim(a)
        if nd < 2:
            raise ValueError(
                f"{self.__name__} requires at least a 2D array with shape (..., vars, obs). "
                "For 1D arrays, use nanvar for variance calculations."
            )
        in_dt = dev.np_dtype_of(a)
        work = kw.loop_dtype(_matrix_loop_dtype(in_dt))
        kw.check_inputs(work, in_dt)
        t = dev.to_device(a, work)
        res = run_matrix(self.__name__, t)
        kw.check_out(work, tuple(res.shape))
        return _finish(res, as_tensor, kw.out)


class ndmovematrix(NumbaBase):
    """numbagg ``ndmovematrix`` (decorators.py:743-818): ``(..., obs, vars) -> (..., obs, vars, vars)``."""

    def __call__(self, a, window: int, min_count: int | None = None, **kwargs):
        kw = _GufuncKwargs(self.__name__, kwargs)
        as_tensor = dev.is_tensor(a)
        if not as_tensor:
            a = np.asarray(a)
        nd = a.dim() if as_tensor else a.ndim
        if nd < 2:
            raise ValueError(f"{self.__name__} requires at least a 2D array with shape (..., obs, vars).")
        if min_count is None:
            min_count = window
        elif min_count < 0:
            raise ValueError(f"min_count must be positive: {min_count}")
        if not 0 < window <= a.shape[-2]:
            raise ValueError(f"window not in valid range: {window}")
        in_dt = dev.np_dtype_of(a)
        work = kw.loop_dtype(_matrix_loop_dtype(in_dt), dtype_ok=(_F64,))  # int64 window / min_count operands
        kw.check_inputs(work, in_dt)
        t = dev.to_device(a, work)
        res = run_matrix(self.__name__, t, window=window, min_count=min_count)
        kw.check_out(work, tuple(res.shape))
        return _finish(res, as_tensor, kw.out)


class ndmoveexpmatrix(NumbaBase):
    """numbagg ``ndmoveexpmatrix`` (decorators.py:1031-1100): one `alpha` per observation
    (a scalar is broadcast), loop dtype from (a, alpha) like NumPy picks it."""

    def __call__(self, a, alpha, min_weight: float = 0, **kwargs):
        kw = _GufuncKwargs(self.__name__, kwargs)
        as_tensor = dev.is_tensor(a)
        if not as_tensor:
            a = np.asarray(a)
        nd = a.dim() if as_tensor else a.ndim
        if nd < 2:
            raise ValueError(f"{self.__name__} requires at least a 2D array with shape (..., obs, vars).")
        n_obs = a.shape[-2]
        if dev.is_tensor(alpha):
            alpha_dt = dev.np_dtype_of(alpha)
        else:
            if not isinstance(alpha, np.ndarray):
                alpha = np.broadcast_to(alpha, n_obs)
            alpha_dt = alpha.dtype
        dts = [dev.np_dtype_of(a), alpha_dt]
        if isinstance(min_weight, np.generic):  # NumPy scalars are strongly typed, Python numbers are not
            dts.append(np.asarray(min_weight).dtype)
        work = kw.loop_dtype(_matrix_loop_dtype(*dts))
        kw.check_inputs(work, dts[0])
        t = dev.to_device(a, work)
        al = dev.to_device(np.ascontiguousarray(alpha) if isinstance(alpha, np.ndarray) else alpha, work, t.device)
        res = run_matrix(self.__name__, t, alpha=al, min_weight=float(min_weight))
        kw.check_out(work, tuple(res.shape))
        return _finish(res, as_tensor, kw.out)
