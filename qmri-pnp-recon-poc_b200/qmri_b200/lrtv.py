"""``x = FISTA_deep(data, param)`` - the reference's LRTV comparison baseline, run on the GPU.

Mirrors ``main_files/algorithms/LRTV/FISTA_deep.m:1-115`` (FISTA on ``0.5 |y - F.forward(x)|^2 + K |x|_TV`` with backtracking)
including the TV prox it calls (``unlocbox/prox/prox_tv.m``), with the argument structs of the call site
``main_recon_tsmis_FFT.m:272-282``:

    data  : {'N', 'M', 'L', 'y', 'F'}            ('D' is unused by the reference)
    param : {'K', 'iter', 'step', 'tol', 'backtrack'}

Returns ``x`` (``N x M x L`` complex double).  ``return_info=True`` also returns ``{'iter', 'step'}``.
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from ._capi import LrtvParams, as_f, check, dtype_code, ptr
from .operators import FOperator


def FISTA_deep(data, param, return_info=False):
    F = data["F"]
    if not isinstance(F, FOperator):
        raise TypeError("data['F'] must come from qmri_b200.fft_operator(P)")
    for k in ("iter", "step", "tol"):
        if k not in param:
            raise KeyError(f"param.{k} is required (FISTA_deep.m:39-48)")
    y = as_f(np.asarray(data["y"]).reshape(-1))
    if not np.iscomplexobj(y):
        y = y.astype(np.complex128)
    if y.size != F.P.nmeas:
        raise ValueError(f"data.y has {y.size} entries, expected {F.P.nmeas}")
    L = int(data.get("L", F.C))
    if (int(data.get("N", F.N)), int(data.get("M", F.M)), L) != (F.N, F.M, F.C):
        raise ValueError(f"data.N/M/L = {data.get('N')}, {data.get('M')}, {L} do not match the operator ({F.N}, {F.M}, {F.C})")
    p = LrtvParams()
    p.K = float(param.get("K", 0.0))
    p.iters = int(param["iter"])
    p.step = float(param["step"])
    p.tol = float(param["tol"])
    p.backtrack = int(bool(param.get("backtrack", 1)))          # FISTA_deep.m:41
    p.tv_tol = float(param.get("tv_tol", 0.0))
    p.tv_maxit = int(param.get("tv_maxit", 0))
    x = np.zeros((F.N, F.M, F.C), np.complex128, order="F")
    its, step = C.c_int(0), C.c_double(0.0)
    check(F.P.ctx.lib.qmri_lrtv(F.P.handle, ptr(y), dtype_code(y), C.byref(p), ptr(x), dtype_code(x), C.byref(its), C.byref(step)))
    return (x, {"iter": its.value, "step": step.value}) if return_info else x
