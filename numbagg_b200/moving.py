"""Moving-window functions (numbagg/moving.py:12-275), computed by nbg_move on the GPU."""

from .decorators import ndmove

move_mean = ndmove("move_mean", doc="NaN-skipping moving mean over a trailing window.")
move_sum = ndmove("move_sum", doc="NaN-skipping moving sum over a trailing window.")
move_std = ndmove("move_std", doc="NaN-skipping moving standard deviation (ddof=1).")
move_var = ndmove("move_var", doc="NaN-skipping moving variance (ddof=1).")
move_cov = ndmove("move_cov", n_inputs=2, doc="Pairwise-complete moving covariance (ddof=1).")
move_corr = ndmove("move_corr", n_inputs=2, doc="Pairwise-complete moving correlation.")

__all__ = ["move_mean", "move_sum", "move_std", "move_var", "move_cov", "move_corr"]
