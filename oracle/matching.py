"""Oracle: MRF dictionary template matching.

Test infrastructure (see ``oracle/__init__.py``).  Restates
``main_files/dictionary_matching/mrf_dtm_cpu.m:38-54,74-98,136-148``:

    x  = single(reshape(data.X,[N,T]))              (mask forced all-true, :51)
    ip = dict.D * ctranspose(x(cind,:))             (:91, K x B)
    [mt, dm] = max(abs(ip),[],1)                    (:92, first index wins ties)
    pd = ip(dm) ./ dict.normD(dm)                   (:94-96)
    qmap = dict.lut(dm,:) ; qmap(isnan(qmap)) = 0   (:138-139)

``precision='f64'`` decides the indices (the truth for parity); ``'f32'`` mirrors
the reference's single-precision arithmetic.  The relative top-2 score gap is
returned so tests can exempt reference near-ties (gap < 1e-6).
"""
from __future__ import annotations

import numpy as np


def mrf_dtm_cpu(dict_, data, par=None, precision="f64", return_gap=False):
    D = np.asarray(dict_["D"])
    normD = np.asarray(dict_["normD"]).reshape(-1)
    lut = np.asarray(dict_["lut"])
    X = np.asarray(data["X"])
    dims = X.shape
    T = dims[-1]
    N = int(np.prod(dims[:-1]))
    Q = lut.shape[1]
    K = D.shape[0]
    if precision == "f64":
        rt, ct = np.float64, np.complex128
    else:
        rt, ct = np.float32, np.complex64
    # reshape(data.X,[N,T]) is column-major; x = single(x) in the reference
    x = X.reshape((N, T), order="F").astype(np.complex64).astype(ct)
    Dm = D.astype(rt)
    blockSize = N
    if par is not None and "fp" in par and "blockSize" in par["fp"]:
        blockSize = int(min(max(np.floor(par["fp"]["blockSize"] / K), 1), N))
    blockSize = min(blockSize, max(1, int(2e8 // K)))  # keep the oracle's K x B block in memory
    mt = np.zeros(N, rt)
    dm = np.zeros(N, np.int64)
    pd = np.zeros(N, ct)
    gap = np.ones(N, np.float64)
    for start in range(0, N, blockSize):
        cind = slice(start, min(start + blockSize, N))
        ip = Dm @ np.conj(x[cind]).T                     # K x B
        a = np.abs(ip)
        d = np.argmax(a, axis=0)                          # first max, like MATLAB
        cols = np.arange(a.shape[1])
        best = a[d, cols]
        mt[cind] = best
        dm[cind] = d
        pd[cind] = ip[d, cols] / normD[d].astype(rt)
        if return_gap:
            a[d, cols] = -1.0
            second = a.max(axis=0) if K > 1 else np.zeros_like(best)
            with np.errstate(divide="ignore", invalid="ignore"):
                g = (best - second) / best
            gap[cind] = np.where(best > 0, g, 0.0)
    qmap = lut[dm].astype(np.float32)
    qmap[np.isnan(qmap)] = 0
    out = {
        "qmap": qmap.reshape(dims[:-1] + (Q,), order="F"),
        "pd": pd.astype(np.complex64).reshape(dims[:-1], order="F"),
        "mt": mt.astype(np.float32).reshape(dims[:-1], order="F"),
        "dm": (dm + 1).reshape(dims[:-1], order="F"),     # 1-based like MATLAB
        "mask": np.ones(dims[:-1], bool),
    }
    if return_gap:
        out["gap"] = gap.reshape(dims[:-1], order="F")
    return out
