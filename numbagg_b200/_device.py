"""Device plumbing: numpy <-> CUDA tensors, (outer, n, inner) views, streams.

PyTorch is used for memory, streams and (in distributed.py) process groups only; all
arithmetic happens in libnbg_b200.so.  numpy inputs are copied to the device and results
copied back (that copy is what bench.py's `e2e` leg times); CUDA tensors are used in place
and a CUDA tensor is returned.
"""

from __future__ import annotations

import math
from typing import Any

import numpy as np
import torch

_NP_TO_TORCH = {
    np.dtype(np.float32): torch.float32,
    np.dtype(np.float64): torch.float64,
    np.dtype(np.float16): torch.float16,
    np.dtype(np.int8): torch.int8,
    np.dtype(np.int16): torch.int16,
    np.dtype(np.int32): torch.int32,
    np.dtype(np.int64): torch.int64,
    np.dtype(np.uint8): torch.uint8,
    np.dtype(np.bool_): torch.bool,
}
_TORCH_TO_NP = {v: k for k, v in _NP_TO_TORCH.items()}
_TORCH_TO_NP[torch.bfloat16] = np.dtype(np.float32)


def require_cuda() -> torch.device:
    if not torch.cuda.is_available():
        raise RuntimeError(
            "numbagg_b200 needs a CUDA device (B200, sm_100a): there is no CPU fallback. "
            "Use numbagg itself on machines without a GPU."
        )
    return torch.device("cuda", torch.cuda.current_device())


def is_tensor(x: Any) -> bool:
    return isinstance(x, torch.Tensor)


def np_dtype_of(x) -> np.dtype:
    """dtype of a tensor / array in NATIVE byte order (NumPy byte-swaps non-native operands on the
    way into a gufunc loop; to_device does the same on the way to the GPU)."""
    if is_tensor(x):
        return _TORCH_TO_NP[x.dtype]
    dt = np.asarray(x).dtype
    return dt if dt.isnative else dt.newbyteorder("=")


def to_device(x, dtype: np.dtype | None = None, device: torch.device | None = None) -> torch.Tensor:
    """numpy / tensor -> CUDA tensor of numpy dtype `dtype` (no copy when already there)."""
    if is_tensor(x):
        t = x
        if not t.is_cuda:
            t = t.to(device or require_cuda(), non_blocking=True)
    else:
        a = np.asarray(x)
        if a.dtype.byteorder == ">" or (a.dtype.byteorder == "=" and not np.little_endian):
            a = a.astype(a.dtype.newbyteorder("="))
        if a.dtype not in _NP_TO_TORCH:
            # uint16/32/64 and friends: widen on the host to a dtype torch can hold
            if a.dtype.kind == "u":
                a = a.astype(np.int64)
            elif a.dtype.kind == "f":
                a = a.astype(np.float64)
            else:
                raise TypeError(f"unsupported dtype {a.dtype}")
        if not a.flags.writeable:
            a = a.copy() if a.size and 0 in a.strides else np.ascontiguousarray(a)
            if not a.flags.writeable:
                a = a.copy()
        if any(s < 0 for s in a.strides):
            a = np.ascontiguousarray(a)
        t = torch.from_numpy(a).to(device or require_cuda(), non_blocking=True)
    if dtype is not None:
        td = _NP_TO_TORCH[np.dtype(dtype)]
        if t.dtype != td:
            t = t.to(td)
    return t


def to_host(t: torch.Tensor, out: np.ndarray | None = None) -> np.ndarray:
    """Device -> host.  Without `out` the result lands in page-locked memory from torch's
    caching host allocator (full-speed D2H; the numpy array keeps the buffer alive)."""
    if out is not None:
        dst = torch.from_numpy(out)
        dst.copy_(t, non_blocking=False)
        return out
    if t.numel() == 0:
        return t.cpu().numpy()
    src = t if t.is_contiguous() else t.contiguous()
    host = torch.empty(src.shape, dtype=src.dtype, pin_memory=True)
    host.copy_(src, non_blocking=True)
    torch.cuda.current_stream().synchronize()
    return host.numpy()


def empty_pinned(shape, dtype) -> np.ndarray:
    """A numpy array backed by page-locked host memory: H2D/D2H copies of such arrays run
    at full PCIe speed and asynchronously (used by bench.py's e2e leg)."""
    t = torch.empty(shape, dtype=_NP_TO_TORCH[np.dtype(dtype)], pin_memory=True)
    return t.numpy()  # the array's base keeps the pinned storage alive


def stream_ptr() -> int:
    return torch.cuda.current_stream().cuda_stream


def ptr(t: torch.Tensor | None) -> int | None:
    if t is None:
        return None
    return t.data_ptr() if t.numel() else None


class CoreView:
    """(outer, n, inner) C-contiguous view of `t` with the core axis in the middle.

    C-contiguous input: pure reshape.  F-contiguous input: the reversed-axes view is
    C-contiguous, so it is used instead (again no copy) and results are permuted back.
    Anything else is made contiguous first (one extra pass; not on any BASELINE config).

    A strided core axis (inner > 1) is handled by the column-walk kernels, one thread per
    column.  When there are too few columns to fill the GPU and the core axis is long (the
    (time, few series) layout with axis=0), the core axis is moved last with one transposing
    copy instead, so the row-tile kernels can parallelise ALONG the core axis.
    """

    MIN_COLUMNS = 32768   # fewer columns than this cannot occupy 148 SMs with one thread each
    MIN_CORE_LEN = 2048

    def __init__(self, t: torch.Tensor, axis: int):
        nd = t.dim()
        if nd == 0:
            raise ValueError("zero-dimensional arrays have no core axis")
        if not -nd <= axis < nd:
            raise np.exceptions.AxisError(axis, nd)
        axis %= nd
        self.transposed = False
        if not t.is_contiguous():
            rev = t.permute(*reversed(range(nd)))
            if rev.is_contiguous():
                t = rev
                axis = nd - 1 - axis
                self.transposed = True
            else:
                t = t.contiguous()
        shape = tuple(t.shape)
        outer = math.prod(shape[:axis])
        inner = math.prod(shape[axis + 1 :])
        self.moved_from = None
        if inner > 1 and outer * inner < self.MIN_COLUMNS and shape[axis] >= self.MIN_CORE_LEN:
            self.moved_from = axis
            t = t.movedim(axis, -1).contiguous()
            axis = nd - 1
            shape = tuple(t.shape)
            outer, inner = outer * inner, 1
        self.t = t
        self.axis = axis
        self.shape = shape
        self.outer = outer
        self.n = shape[axis]
        self.inner = inner

    def like(self, other: torch.Tensor) -> torch.Tensor:
        """Bring an operand shaped like the data (along all but possibly the core axis) into
        this view's memory order."""
        if self.transposed:
            other = other.permute(*reversed(range(other.dim())))
        if self.moved_from is not None:
            other = other.movedim(self.moved_from, -1)
        return other if other.is_contiguous() else other.contiguous()

    def restore(self, out: torch.Tensor) -> torch.Tensor:
        if self.moved_from is not None:
            out = out.movedim(-1, self.moved_from)
        if self.transposed:
            return out.permute(*reversed(range(out.dim())))
        return out
