"""numbagg_b200 -- B200-native (sm_100a) implementation of numbagg's data-parallel hot path.

Drop-in for the moving-window, exponential-moving, grouped, fill and plain NaN-reduction
functions of numbagg
(same names, keyword-only parameters, axis semantics and validation errors; see
numbagg/__init__.py:3-62).  Everything is computed by hand-written CUDA kernels in
libnbg_b200.so through the C ABI in include/nbg_b200.h; there is no CPU fallback.
"""

from ._device import empty_pinned
from ._lib import LIB_PATH, NbgError, launch_count
from .funcs import (
    allnan,
    anynan,
    bfill,
    count,
    ffill,
    nanargmax,
    nanargmin,
    nancount,
    nanmax,
    nanmean,
    nanmedian,
    nanmin,
    nanquantile,
    nanstd,
    nansum,
    nanvar,
)
from .grouped import (
    group_nanall,
    group_nanany,
    group_nanargmax,
    group_nanargmin,
    group_nancount,
    group_nanfirst,
    group_nanlast,
    group_nanmax,
    group_nanmean,
    group_nanmin,
    group_nanprod,
    group_nanstd,
    group_nansum,
    group_nansum_of_squares,
    group_nanvar,
)
from .moving import move_corr, move_cov, move_mean, move_std, move_sum, move_var
from .moving_matrix import (
    move_corrmatrix,
    move_covmatrix,
    move_exp_nancorrmatrix,
    move_exp_nancovmatrix,
    nancorrmatrix,
    nancovmatrix,
)
from .moving_exp import (
    move_exp_nancorr,
    move_exp_nancount,
    move_exp_nancov,
    move_exp_nanmean,
    move_exp_nanstd,
    move_exp_nansum,
    move_exp_nanvar,
)

GROUPED_FUNCS = [
    group_nanall, group_nanany, group_nanargmax, group_nanargmin, group_nancount, group_nanfirst,
    group_nanlast, group_nanmax, group_nanmean, group_nanmin, group_nanprod, group_nanstd,
    group_nansum, group_nansum_of_squares, group_nanvar,
]
MOVE_EXP_FUNCS = [
    move_exp_nancorr, move_exp_nancount, move_exp_nancov, move_exp_nanmean, move_exp_nanstd,
    move_exp_nansum, move_exp_nanvar,
]
MOVE_FUNCS = [move_corr, move_cov, move_mean, move_std, move_sum, move_var]
OTHER_FUNCS = [bfill, ffill]
AGGREGATION_FUNCS = [
    allnan, anynan, nancount, nansum, nanmean, nanvar, nanstd, nanargmax, nanargmin, nanmax, nanmin,
]
QUANTILE_FUNCS = [nanquantile, nanmedian]
MATRIX_FUNCS = [
    nancorrmatrix, nancovmatrix, move_corrmatrix, move_covmatrix, move_exp_nancorrmatrix, move_exp_nancovmatrix,
]

__version__ = "0.1.0"

__all__ = [
    *(f.__name__ for f in GROUPED_FUNCS + MOVE_EXP_FUNCS + MOVE_FUNCS + OTHER_FUNCS + AGGREGATION_FUNCS),
    "count", "nanquantile", "nanmedian", *(f.__name__ for f in MATRIX_FUNCS), "MATRIX_FUNCS", "AGGREGATION_FUNCS", "QUANTILE_FUNCS", "GROUPED_FUNCS", "MOVE_EXP_FUNCS", "MOVE_FUNCS", "OTHER_FUNCS",
    "empty_pinned", "launch_count", "NbgError", "LIB_PATH",
]
