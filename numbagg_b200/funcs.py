"""ffill / bfill (numbagg/funcs.py:294-326), computed by nbg_fill on the GPU."""

from .decorators import ndfill

ffill = ndfill("ffill", doc="Forward fill missing values.")
bfill = ndfill("bfill", doc="Backward fill missing values.")

__all__ = ["ffill", "bfill"]
