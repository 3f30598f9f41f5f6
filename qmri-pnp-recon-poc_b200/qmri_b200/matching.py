"""``out = mrf_dtm_cpu(dict, data, par)`` - dictionary template matching, run on the GPU.

The name is the reference's (``main_files/dictionary_matching/mrf_dtm_cpu.m:1-166``) so that
``main_recon_tsmis_FFT.m:317`` keeps working unchanged; ``mrf_dtm`` is an alias.

    dict  : {'D': [K x C] unit-norm atoms, 'normD': [K], 'lut': [K x Q]}            (:8-12)
    data  : {'X': [Nx, Ny, (Nz,) T]}   ('mask' is ignored: forced all-true, :51)
    par   : {'f': {'qout','pdout','mtout','dmout','Xout','Yout','verbose'}, 'fp': {'blockSize'}}
    out   : {'qmap','mask'} | {'pd'} | {'mt'} | {'dm'} | {'Xfit','X'} | {'Y'}         (:126-164)

``par.fp.blockSize`` only bounds the reference's materialised K x B score block (:74-78); the
fused kernel never materialises scores, so it is accepted and has no effect on the result.
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import _capi
from ._capi import Context, as_f, check, dtype_code, ptr


class Dictionary:
    """Device-resident ``dict`` (qmri_dict).  ``shard`` = (begin, end) atom range scored here."""

    def __init__(self, dict_, ctx=None, shard=None, shard_only=False):
        """``shard_only=True``: ``dict_['D']`` holds ONLY the atoms ``[shard[0], shard[1])`` (``normD`` / ``lut`` stay whole,
        12 B per atom) - per-rank memory and upload time scale with K / world (``qmri_dict_load_shard``)."""
        self.ctx = ctx or Context.default()
        D = np.asfortranarray(np.asarray(dict_["D"]))
        if np.iscomplexobj(D):
            if np.abs(D.imag).max() > 0:
                raise _capi.QmriError(_capi.QMRI_EUNSUPPORTED, "complex dictionaries are out of scope (the reference uses the real_ dictionaries)")
            D = D.real
        D = np.asfortranarray(D.astype(np.float32))
        if D.ndim != 2:
            raise ValueError("dict.D must be K x C")
        self.K, self.C = D.shape
        normD = np.ascontiguousarray(np.asarray(dict_["normD"], dtype=np.float32).reshape(-1))
        self.shard_only = bool(shard_only)
        if self.shard_only:
            if shard is None:
                raise ValueError("shard_only needs shard=(begin, end)")
            if D.shape[0] != int(shard[1]) - int(shard[0]):
                raise ValueError(f"shard_only: dict.D must hold the {int(shard[1]) - int(shard[0])} atoms of the shard, got {D.shape[0]} rows")
            self.K = normD.size
        lut = np.asfortranarray(np.asarray(dict_["lut"], dtype=np.float32))
        if lut.ndim != 2 or lut.shape[0] != self.K or normD.size != self.K:
            raise ValueError("dict.lut must be K x Q and dict.normD must have K entries")
        self.Q = lut.shape[1]
        self.host_D, self.host_normD = D, normD
        self.shard = (0, self.K) if shard is None else (int(shard[0]), int(shard[1]))
        h = C.c_void_p()
        load = self.ctx.lib.qmri_dict_load_shard if self.shard_only else self.ctx.lib.qmri_dict_load
        check(load(self.ctx.handle, ptr(D), ptr(normD), ptr(lut), self.K, self.C, self.Q, self.shard[0], self.shard[1], C.byref(h)))
        self.handle = h

    def close(self):
        if self.handle:
            self.ctx.lib.qmri_dict_destroy(self.handle)
            self.handle = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def _flag(par, name, default=0):
    return bool(par.get("f", {}).get(name, default)) if par else bool(default)


def mrf_dtm_cpu(dict_, data, par=None):
    own = not isinstance(dict_, Dictionary)
    d = Dictionary(dict_) if own else dict_
    try:
        X = np.asarray(data["X"])
        if X.ndim not in (3, 4):
            raise ValueError("data.X must be [Nx, Ny, T] or [Nx, Ny, Nz, T]")
        dims = X.shape
        T = dims[-1]
        if T != d.C:
            raise ValueError(f"data.X has {T} timepoints but dict.D has {d.C} columns")
        npix = int(np.prod(dims[:-1]))
        # x = single(reshape(data.X,[N,T])) (:50-54)
        x = as_f(X).reshape((npix, T), order="F")
        x = np.asfortranarray(x.astype(np.complex64 if np.iscomplexobj(x) else np.float32))
        want_q, want_pd = _flag(par, "qout", 1), _flag(par, "pdout", 1)
        want_mt, want_dm, want_X = _flag(par, "mtout"), _flag(par, "dmout"), _flag(par, "Xout")
        if par and par.get("f", {}).get("verbose"):
            print("Matching data ")
        qmap = np.zeros((npix, d.Q), np.float32, order="F") if want_q else None
        pd = np.zeros(npix, np.complex64) if (want_pd or want_X) else None
        mt = np.zeros(npix, np.float32) if want_mt else None
        dm = np.zeros(npix, np.int32) if (want_dm or want_X) else None
        check(d.ctx.lib.qmri_match(d.handle, ptr(x), dtype_code(x), npix, ptr(qmap), ptr(pd), ptr(mt), ptr(dm)))
        out = {}
        sp = dims[:-1]
        if want_X:  # X(fit) = ip(dm) .* D(dm,:)   (:95, :129-134) - output formatting from the kernel's results
            Dh = np.asarray(dict_["D"] if own else dict_.host_D, dtype=np.float32)
            nD = np.asarray(dict_["normD"] if own else dict_.host_normD, dtype=np.float32).reshape(-1)
            Xfit = (pd * nD[dm - 1])[:, None] * Dh[dm - 1]
            out["Xfit"] = Xfit.astype(np.complex64 if np.iscomplexobj(X) else np.float32).reshape(dims, order="F")
            out["X"] = data["X"]
        if want_q:
            out["qmap"] = qmap.reshape(sp + (d.Q,), order="F")
            out["mask"] = np.ones(sp, dtype=bool)
        if want_pd:
            out["pd"] = (pd if np.iscomplexobj(X) else pd.real.copy()).reshape(sp, order="F")
        if want_mt:
            out["mt"] = mt.reshape(sp, order="F")
        if want_dm:
            out["dm"] = dm.astype(np.float32).reshape(sp, order="F")  # single(dm), :158
        if _flag(par, "Yout") and "Y" in data:
            out["Y"] = data["Y"]
        return out
    finally:
        if own:
            d.close()


mrf_dtm = mrf_dtm_cpu


def synthesize_tsmis(dict_, qmap_slice, return_index=False):
    """``main_synthesize_tsmis.m:84-98`` for one slice: ``qmap_slice`` ``[3 x N x M]`` (T1, T2, PD) -> X ``[N x M x C]`` real
    single (what the script saves); ``return_index`` also returns the 0-based nearest atom per pixel (pixel n + N m)."""
    own = not isinstance(dict_, Dictionary)
    d = Dictionary(dict_) if own else dict_
    try:
        q3 = np.asarray(qmap_slice)
        if q3.ndim != 3 or q3.shape[0] != 3:
            raise ValueError("qmap slice must be [3 x N x M] (T1, T2, PD)")
        _, N, M = q3.shape
        qm = np.asfortranarray(np.transpose(q3, (1, 2, 0)).reshape((-1, 3), order="F").astype(np.float32))  # :84-87
        npix = qm.shape[0]
        X = np.zeros((npix, d.C), np.float32, order="F")
        I = np.zeros(npix, np.int32)
        check(d.ctx.lib.qmri_synthesize(d.handle, ptr(qm), npix, ptr(X), ptr(I)))
        X = X.reshape((N, M, d.C), order="F")
        return (X, I.astype(np.int64) - 1) if return_index else X
    finally:
        if own:
            d.close()
