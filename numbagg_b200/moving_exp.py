"""Exponentially-weighted moving functions (numbagg/moving_exp.py:12-335), computed by
nbg_move_exp on the GPU."""

from .decorators import ndmoveexp

move_exp_nancount = ndmoveexp("move_exp_nancount", doc="Exponentially decayed count of valid values.")
move_exp_nanmean = ndmoveexp("move_exp_nanmean", doc="Exponentially weighted moving mean")
move_exp_nansum = ndmoveexp("move_exp_nansum", doc="Exponentially decayed moving sum.")
move_exp_nanvar = ndmoveexp("move_exp_nanvar", doc="Exponentially weighted, bias-corrected moving variance.")
move_exp_nanstd = ndmoveexp("move_exp_nanstd", doc="Square root of the exponentially weighted moving variance.")
move_exp_nancov = ndmoveexp("move_exp_nancov", n_inputs=2, doc="Exponentially weighted moving covariance.")
move_exp_nancorr = ndmoveexp("move_exp_nancorr", n_inputs=2, doc="Exponentially weighted moving correlation.")

__all__ = [
    "move_exp_nancount", "move_exp_nanmean", "move_exp_nansum", "move_exp_nanvar",
    "move_exp_nanstd", "move_exp_nancov", "move_exp_nancorr",
]
