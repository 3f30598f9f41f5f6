"""Oracle: the PnP-ADMM loop.

Test infrastructure (see ``oracle/__init__.py``).  Faithful restatement of
``main_files/algorithms/PnP_ADMM/PnP_ADMM.m:76-78,93-146``: state ``x`` complex,
``v`` (complex at iteration 1 because ``v = x = X0``, real afterwards) and
``uold`` (scalar 0 at iteration 1, complex array afterwards).  The denoiser is a
callback taking the N x M x Cin array in [0,1] (``PnP_ADMM.m:128,132-133``).
"""
from __future__ import annotations

import numpy as np

from .sampling import FOperator
from .xupdate import (norm_zero_to_one, undo_norm_zero_to_one, xupdate_closed_form,
                      xupdate_exact, xupdate_lsqr)


def pnp_admm(y, param, solver="exact", trace=None):
    """``x = PnP_ADMM(y, param)`` for one slice (N x M x C).

    ``param`` keys follow the reference struct (``main_recon_tsmis_FFT.m:164-170,
    285-292``): ``iter, gamma, F, cg_tol, X0, net, denoiser_type, noise_map``.
    ``solver``: 'exact' | 'closed' | 'lsqr' selects the x-update formulation.
    ``trace``: optional list receiving per-iteration dicts (x, v_in, v, u).
    """
    max_iter = int(param["iter"])
    r = float(param["gamma"])
    F: FOperator = param["F"]
    net = param["net"]
    dtype = param.get("denoiser_type", "single_level")
    x = np.asarray(param["X0"], dtype=np.complex128)
    v = x
    uold = 0.0
    for _ in range(max_iter):
        z = v - uold
        if solver == "exact":
            x = xupdate_exact(F, y, z, r)
        elif solver == "closed":
            x = xupdate_closed_form(F, y, z, r)
        else:
            x, _its = xupdate_lsqr(F, y, z, r, tol=float(param.get("cg_tol", 1e-4)), maxit=100, x0=x)
        v = np.real(x + uold)                                   # :115-118
        v_in, x_min, x_max, x_range = norm_zero_to_one(v)       # :121
        if dtype == "multi_level":
            v_in = np.concatenate([v_in, np.asarray(param["noise_map"])[:, :, None]], axis=2)  # :132
        v = np.asarray(net(v_in), dtype=np.float64)             # :128/:133
        v = undo_norm_zero_to_one(v, x_min, x_max, x_range)     # :138
        uold = uold + x - v                                     # :144
        if trace is not None:
            trace.append({"x": x, "v_in": v_in, "v": v, "u": uold})
    return x


def pnp_admm_wv(y, param):
    """The (w, v) state reformulation, valid when A A^H = I.

    w_k = x_k + u_{k-1}.  One step: z = 2 v - w, x = z + A^H(y - A z)/(1+rho),
    w' = x + (w - v).  Iteration 1 is a no-op on x (A X0 = y), so w_1 = X0.
    Returns the final x, like ``PnP_ADMM``.
    """
    max_iter = int(param["iter"])
    r = float(param["gamma"])
    F: FOperator = param["F"]
    net = param["net"]
    w = np.asarray(param["X0"], dtype=np.complex128)
    x = w
    for k in range(max_iter):
        if k > 0:
            z = 2.0 * v - w
            x = xupdate_closed_form(F, y, z, r)
            w = x + (w - v)
        v_in, x_min, x_max, x_range = norm_zero_to_one(np.real(w))
        v = undo_norm_zero_to_one(np.asarray(net(v_in), dtype=np.float64), x_min, x_max, x_range)
    return x
