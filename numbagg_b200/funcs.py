"""Plain functions of numbagg/funcs.py on the GPU: ffill / bfill (:294-326, nbg_fill) and the
NaN-aware reductions (:23-242, nbg_reduce) and nanquantile / nanmedian (:245-291, 332-335,
nbg_quantile)."""

from .decorators import ndaggregate, ndfill, ndquantile, ndreduce

ffill = ndfill("ffill", doc="Forward fill missing values.")
bfill = ndfill("bfill", doc="Backward fill missing values.")

allnan = ndaggregate("allnan", doc="True where every element along `axis` is NaN.")
anynan = ndaggregate("anynan", doc="True where any element along `axis` is NaN.")
nancount = ndaggregate("nancount", doc="Number of non-NaN elements along `axis`.")
nansum = ndaggregate("nansum", doc="Sum of the non-NaN elements along `axis`.")
nanmean = ndaggregate("nanmean", doc="Mean of the non-NaN elements along `axis`.")
nanvar = ndaggregate("nanvar", supports_ddof=True, doc="Variance of the non-NaN elements along `axis`.")
nanstd = ndaggregate("nanstd", supports_ddof=True, doc="Standard deviation of the non-NaN elements along `axis`.")
count = nancount  # numbagg/funcs.py:329
nanargmax = ndreduce("nanargmax", doc="Flat index of the first maximum, ignoring NaN.")
nanargmin = ndreduce("nanargmin", doc="Flat index of the first minimum, ignoring NaN.")
nanmax = ndreduce("nanmax", doc="Maximum, ignoring NaN.")
nanmin = ndreduce("nanmin", doc="Minimum, ignoring NaN.")

nanquantile = ndquantile("nanquantile", doc="Quantiles of the non-NaN elements along `axis` (linear interpolation).")


def nanmedian(a, *, axis=None, **kwargs):
    """Median of the non-NaN elements (numbagg/funcs.py:332-335)."""
    return nanquantile(a, quantiles=0.5, axis=axis, **kwargs)


__all__ = [
    "ffill", "bfill", "allnan", "anynan", "nancount", "count", "nansum", "nanmean", "nanvar", "nanstd",
    "nanargmax", "nanargmin", "nanmax", "nanmin", "nanquantile", "nanmedian",
]
