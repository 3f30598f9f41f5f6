"""Foreground mask and quality metrics of the reference's driver script, computed on the GPU.

Mirrors (same names and argument meaning):

* ``mask = getmask_fromPD(PD, thresh)`` - ``main_files/utils/getmask_fromPD.m:1-15``
  (``main_recon_tsmis_FFT.m:190``: ``foreground_mask = getmask_fromPD(qmap0(:,:,3), 0.15)``).
* ``recon_metrics(qmap, qmap0, foreground_mask, X, X0)`` - the block ``main_recon_tsmis_FFT.m:328-384``:
  masked MAE, ``psnr`` and ``ssim`` of T1 / T2 / PD and the mean ``psnr`` / ``ssim`` of ``abs`` of the TSMI channels,
  returned under the script's variable names.
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from ._capi import Context, as_f, check, dtype_code, ptr

METRIC_NAMES = ("tsmi_mean_psnr", "tsmi_mean_ssim", "t1_mae", "t1_psnr", "t1_ssim", "t2_mae", "t2_psnr", "t2_ssim",
                "pd_mae", "pd_psnr", "pd_ssim")


def getmask_fromPD(PD, thresh, ctx=None):
    ctx = ctx or Context.default()
    PD = as_f(PD)
    if PD.ndim != 2:
        raise ValueError("PD must be a 2-D image")
    N, M = PD.shape
    mask = np.zeros((N, M), np.float32, order="F")
    check(ctx.lib.qmri_foreground_mask(ctx.handle, ptr(PD), dtype_code(PD), N, M, float(thresh), ptr(mask)))
    return mask.astype(np.float64)   # the reference's mask is a double image of 0 / 1


def recon_metrics(qmap, qmap0, foreground_mask=None, X=None, X0=None, ctx=None):
    """``qmap = cat(3, out.qmap, out.pd)`` (N x M x 3), ``qmap0`` the ground truth; ``X`` / ``X0`` optional TSMIs."""
    ctx = ctx or Context.default()
    qmap, qmap0 = as_f(qmap), as_f(qmap0)
    if qmap.ndim != 3 or qmap.shape[2] != 3 or qmap0.shape != qmap.shape:
        raise ValueError("qmap and qmap0 must both be N x M x 3 (T1, T2, PD)")
    N, M, _ = qmap.shape
    mask = None
    if foreground_mask is not None:
        mask = np.asfortranarray(np.asarray(foreground_mask, dtype=np.float32))
        if mask.shape != (N, M):
            raise ValueError(f"foreground_mask must be {N} x {M}")
    Cc = 0
    if (X is None) != (X0 is None):
        raise ValueError("pass both X and X0 or neither")
    if X is not None:
        X, X0 = as_f(X), as_f(X0)
        if X.ndim != 3 or X.shape[:2] != (N, M) or X0.shape != X.shape:
            raise ValueError("X and X0 must both be N x M x C")
        Cc = X.shape[2]
    out = np.zeros(11, np.float64)
    check(ctx.lib.qmri_recon_metrics(ctx.handle, N, M, Cc, ptr(qmap), dtype_code(qmap), ptr(qmap0), dtype_code(qmap0),
                                     ptr(mask), ptr(X), dtype_code(X) if X is not None else 0,
                                     ptr(X0), dtype_code(X0) if X0 is not None else 0, ptr(out)))
    return dict(zip(METRIC_NAMES, (float(v) for v in out)))
