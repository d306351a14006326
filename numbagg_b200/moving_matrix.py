"""Matrix functions of numbagg on the GPU (nbg_matrix): the static pairwise-complete
covariance / correlation matrices (numbagg/funcs.py:338-532) and their moving-window and
exponentially weighted forms (numbagg/moving_matrix.py:16-432)."""

from .decorators import ndmatrix, ndmoveexpmatrix, ndmovematrix

nancorrmatrix = ndmatrix("nancorrmatrix", doc="Correlation matrix of (..., vars, obs), pairwise-complete observations.")
nancovmatrix = ndmatrix("nancovmatrix", doc="Covariance matrix of (..., vars, obs), pairwise-complete observations.")
move_corrmatrix = ndmovematrix("move_corrmatrix", doc="Moving-window correlation matrices of (..., obs, vars).")
move_covmatrix = ndmovematrix("move_covmatrix", doc="Moving-window covariance matrices of (..., obs, vars).")
move_exp_nancorrmatrix = ndmoveexpmatrix("move_exp_nancorrmatrix", doc="Exponentially weighted correlation matrices.")
move_exp_nancovmatrix = ndmoveexpmatrix("move_exp_nancovmatrix", doc="Exponentially weighted covariance matrices.")

__all__ = ["nancorrmatrix", "nancovmatrix", "move_corrmatrix", "move_covmatrix", "move_exp_nancorrmatrix",
           "move_exp_nancovmatrix"]
