"""Acquisition operator: the reference's `setup_subsampling_*` and `F.forward / F.adjoint`.

Mirrors (same names, argument meaning and error behaviour):

* ``setup_subsampling_spiralgrided(N, M, S, V)`` -
  ``main_files/subsampling_patterns/setup_subsampling_spiralgrided.m:1-43``
* ``setup_subsampling_epi(N, M, percentage, V)`` -
  ``main_files/subsampling_patterns/setup_subsampling_epi.m:1-37``
* ``F.forward = @(x) P.for(reshape(fft2(x),[],1))/sqrt(N*M)`` and
  ``F.adjoint = @(x) ifft2(reshape(P.adj(x),N,M,[]))*sqrt(N*M)`` -
  ``main_recon_tsmis_FFT.m:228-229``  ->  ``fft_operator(P)``.

Arrays use the MATLAB shapes: images ``N x M x C`` (optionally ``x S`` slices as a
trailing axis = S independent reference calls), measurements ``nmeas`` (``x S``).
All arithmetic happens on the GPU through the C ABI.
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import _capi
from ._capi import Context, as_f, check, dtype_code, ptr


class SubsamplingPattern:
    """What ``setup_subsampling_*`` returns (the struct the reference calls P / K / S)."""

    def __init__(self, ctx, handle, N, M, C_, L):
        self.ctx, self.handle = ctx, handle
        self.N, self.M, self.C, self.L = N, M, C_, L
        self.nmeas = int(ctx.lib.qmri_op_nmeas(handle))

    def indices(self):
        """0-based ascending column-major k indices per frame (MATLAB ``find`` minus one)."""
        idx = np.zeros(self.nmeas, np.int32)
        fp = np.zeros(self.L + 1, np.int64)
        check(self.ctx.lib.qmri_op_indices(self.handle, ptr(idx), ptr(fp)))
        return idx, fp

    # P.for / P.adj, the handles the reference constructors return (setup_subsampling_spiralgrided.m:41-42,
    # setup_subsampling_epi.m:34-35).  `for` is a Python keyword: the forward handle is `P.for_` (also P["for"] / P["adj"]).
    def for_(self, vec):
        """``y = P.for(vec)``: ``vec`` = the N*M*C k-space column ``reshape(fft2(x),[],1)`` -> ``nmeas`` samples."""
        v = as_f(np.asarray(vec).reshape(-1, order="F"))
        if v.size != self.N * self.M * self.C:
            raise ValueError(f"P.for expects {self.N * self.M * self.C} entries, got {v.size}")
        y = np.zeros(self.nmeas, np.complex64 if v.dtype in (np.float32, np.complex64) else np.complex128)
        check(self.ctx.lib.qmri_op_for(self.handle, ptr(v), dtype_code(v), ptr(y), dtype_code(y)))
        return y

    def adj(self, y):
        """``vec = P.adj(y)`` = ``P' * y``: ``nmeas`` samples -> the N*M*C k-space column."""
        y = as_f(np.asarray(y).reshape(-1, order="F"))
        if y.size != self.nmeas:
            raise ValueError(f"P.adj expects {self.nmeas} entries, got {y.size}")
        v = np.zeros(self.N * self.M * self.C, np.complex64 if y.dtype in (np.float32, np.complex64) else np.complex128)
        check(self.ctx.lib.qmri_op_adj(self.handle, ptr(y), dtype_code(y), ptr(v), dtype_code(v)))
        return v

    def __getitem__(self, name):
        if name == "for":
            return self.for_
        if name == "adj":
            return self.adj
        raise KeyError(name)

    def close(self):
        if self.handle:
            self.ctx.lib.qmri_op_destroy(self.handle)
            self.handle = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def _prep_V(V):
    V = np.atleast_2d(np.asarray(V))
    if np.iscomplexobj(V):
        if np.abs(V.imag).max() > 0:
            raise _capi.QmriError(_capi.QMRI_EUNSUPPORTED, "V must be real (the reference passes real(dict.V))")
        V = V.real
    return np.asfortranarray(V.astype(np.float64))


def setup_subsampling_spiralgrided(N, M, S, V, ctx=None):
    ctx = ctx or Context.default()
    V = _prep_V(V)
    L, Cc = V.shape
    h = C.c_void_p()
    check(ctx.lib.qmri_op_spiral(ctx.handle, int(N), int(M), int(S), ptr(V), L, Cc, C.byref(h)))
    return SubsamplingPattern(ctx, h, int(N), int(M), Cc, L)


def setup_subsampling_epi(N, M, percentage, V, ctx=None):
    ctx = ctx or Context.default()
    V = _prep_V(V)
    L, Cc = V.shape
    h = C.c_void_p()
    check(ctx.lib.qmri_op_epi(ctx.handle, int(N), int(M), float(percentage), ptr(V), L, Cc, C.byref(h)))
    return SubsamplingPattern(ctx, h, int(N), int(M), Cc, L)


def setup_subsampling_explicit(N, M, frames, V, ctx=None):
    """Operator from explicit per-frame index lists (0-based, ascending) - `qmri_op_create`."""
    ctx = ctx or Context.default()
    V = _prep_V(V)
    L, Cc = V.shape
    idx = np.concatenate([np.asarray(f, np.int32) for f in frames]) if len(frames) else np.zeros(0, np.int32)
    fp = np.concatenate([[0], np.cumsum([len(f) for f in frames])]).astype(np.int64)
    h = C.c_void_p()
    check(ctx.lib.qmri_op_create(ctx.handle, int(N), int(M), Cc, L, ptr(np.ascontiguousarray(idx)), ptr(fp), ptr(V), C.byref(h)))
    return SubsamplingPattern(ctx, h, int(N), int(M), Cc, L)


class FOperator:
    """``F.forward`` / ``F.adjoint`` (``main_recon_tsmis_FFT.m:228-229``) on the GPU."""

    def __init__(self, P: SubsamplingPattern):
        self.P = P
        self.N, self.M, self.C = P.N, P.M, P.C

    def _slices(self, x, nd):
        if x.ndim == nd:
            return 1, False
        if x.ndim == nd + 1:
            return x.shape[-1], True
        raise ValueError(f"expected {nd} or {nd + 1} dimensions, got {x.ndim}")

    def forward(self, x):
        x = as_f(x)
        S, batched = self._slices(x, 3)
        if x.shape[:3] != (self.N, self.M, self.C):
            raise ValueError(f"x must be {self.N} x {self.M} x {self.C}, got {x.shape}")
        cd = np.complex64 if x.dtype in (np.float32, np.complex64) else np.complex128
        y = np.zeros((self.P.nmeas, S), cd, order="F")
        check(self.P.ctx.lib.qmri_forward(self.P.handle, ptr(x), dtype_code(x), S, ptr(y), dtype_code(y)))
        return y if batched else y[:, 0]

    def adjoint(self, y):
        y = as_f(y)
        if not np.iscomplexobj(y):
            y = y.astype(np.complex128)
        S, batched = self._slices(y, 1)
        if y.shape[0] != self.P.nmeas:
            raise ValueError(f"y must have {self.P.nmeas} rows, got {y.shape}")
        x = np.zeros((self.N, self.M, self.C, S), y.dtype, order="F")
        check(self.P.ctx.lib.qmri_adjoint(self.P.handle, ptr(y), dtype_code(y), S, ptr(x), dtype_code(x)))
        return x if batched else x[..., 0]

    def xupdate(self, y, v, u, rho, want_w=False):
        """One exact least-squares step (``PnP_ADMM.m:102``): returns x [, w = x+u, (min,max) of real(w)]."""
        y = as_f(y)
        if not np.iscomplexobj(y):
            y = y.astype(np.complex128)
        v = as_f(v)
        S, batched = self._slices(v, 3)
        uu = None
        if u is not None and not np.isscalar(u):
            uu = as_f(u)
        elif u is not None and u != 0:
            uu = as_f(np.full(v.shape, u, dtype=np.complex128))
        x = np.zeros((self.N, self.M, self.C, S), np.complex128, order="F")
        w = np.zeros_like(x, order="F") if want_w else None
        mm = np.zeros((S, 2), np.float32) if want_w else None
        check(self.P.ctx.lib.qmri_xupdate(self.P.handle, float(rho), ptr(y), dtype_code(y), ptr(v), dtype_code(v),
                                          ptr(uu), dtype_code(uu) if uu is not None else 0, S, ptr(x), dtype_code(x),
                                          ptr(w), dtype_code(w) if w is not None else 0, ptr(mm)))
        if not batched:
            x = x[..., 0]
            if want_w:
                w, mm = w[..., 0], mm[0]
        return (x, w, mm) if want_w else x


def fft_operator(P):
    return FOperator(P)


def awgn(Y, snr, mode="measured", seed=0, ctx=None):
    """``Y = awgn(Y, snr, 'measured')`` (``main_recon_tsmis_FFT.m:243``) on the GPU: complex white Gaussian noise of power
    ``mean(|Y|^2) / 10^(snr/10)``, per slice (column) of ``Y``.  Philox-4x32-10 keyed by ``seed``: reproducible, but not
    MATLAB's random stream (statistical equality only)."""
    if mode != "measured":
        raise ValueError("only awgn(Y, snr, 'measured') is used by the reference (main_recon_tsmis_FFT.m:243)")
    ctx = ctx or Context.default()
    Y = np.array(Y, order="F", copy=True)
    if not np.iscomplexobj(Y):
        Y = Y.astype(np.complex128)
    if Y.dtype not in (np.complex64, np.complex128):
        Y = Y.astype(np.complex128)
    if Y.ndim not in (1, 2):
        raise ValueError("Y must be nmeas or nmeas x S")
    nmeas = Y.shape[0]
    S = Y.shape[1] if Y.ndim == 2 else 1
    check(ctx.lib.qmri_awgn(ctx.handle, ptr(Y), dtype_code(Y), nmeas, S, float(snr), int(seed) & 0xFFFFFFFFFFFFFFFF))
    return Y
