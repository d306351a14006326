"""Grouped NaN-aware reductions (numbagg/grouped.py:7-270), computed by nbg_group on the GPU."""

from .decorators import groupndreduce

group_nanmean = groupndreduce("group_nanmean", supports_ints=False)
group_nansum = groupndreduce("group_nansum")
group_nancount = groupndreduce("group_nancount")
group_nanargmax = groupndreduce("group_nanargmax")
group_nanargmin = groupndreduce("group_nanargmin")
group_nanfirst = groupndreduce("group_nanfirst")
group_nanlast = groupndreduce("group_nanlast")
group_nanprod = groupndreduce("group_nanprod")
group_nansum_of_squares = groupndreduce("group_nansum_of_squares")
group_nanvar = groupndreduce("group_nanvar", supports_bool=False, supports_ints=False, supports_ddof=True)
group_nanstd = groupndreduce("group_nanstd", supports_bool=False, supports_ints=False, supports_ddof=True)
group_nanmin = groupndreduce("group_nanmin")
group_nanmax = groupndreduce("group_nanmax")
group_nanany = groupndreduce("group_nanany")
group_nanall = groupndreduce("group_nanall")

__all__ = [
    "group_nanmean", "group_nansum", "group_nancount", "group_nanargmax", "group_nanargmin",
    "group_nanfirst", "group_nanlast", "group_nanprod", "group_nansum_of_squares", "group_nanvar",
    "group_nanstd", "group_nanmin", "group_nanmax", "group_nanany", "group_nanall",
]
