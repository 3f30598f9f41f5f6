"""Oracle cross-check: a deliberately LITERAL transcription of the reference's sparse-matrix operator and lsqr call.

Test infrastructure (see ``oracle/__init__.py``).  ``oracle/sampling.py`` / ``oracle/xupdate.py`` apply ``P`` matrix-free
and solve the x-update in closed form; this module instead builds exactly what the MATLAB code builds - the explicit
sparse matrix

    tmp = sparse([1:numel(ind)], ind, ones(1,numel(ind)), numel(ind), N*M);
    P   = [P; tmp * kron(conj(V(i,:)), speye(N*M))];
        (setup_subsampling_spiralgrided.m:36-37, setup_subsampling_epi.m:31-32)

with ``scipy.sparse`` (1-based index arithmetic kept, converted at the last moment), the closures
``F.forward = @(x) P.for(reshape(fft2(x),[],1))/sqrt(N*M)``, ``F.adjoint = @(x) ifft2(reshape(P.adj(x),N,M,[]))*sqrt(N*M)``
(``main_recon_tsmis_FFT.m:228-229``) and the solver call of ``PnP_ADMM.m:102`` through ``scipy.sparse.linalg.lsqr`` on the
stacked operator of ``afun`` (``PnP_ADMM.m:153-171``).  It shares no code with the other two modules, so agreement
between them (tests/test_oracle_literal.py) is an independent check of the restatement - the strongest pin available
without MATLAB / Octave (parity of the MATLAB half stays "unpinned by the reference": it holds no vectors).
"""
from __future__ import annotations

import numpy as np
import scipy.sparse as sp
from scipy.sparse.linalg import LinearOperator, lsqr


def _round_half_away(x):
    return np.where(x >= 0, np.floor(x + 0.5), -np.floor(-x + 0.5))


def spiral_P(N, M, S, V):
    """setup_subsampling_spiralgrided.m:7-42, line by line (1-based indices as in MATLAB)."""
    V = np.atleast_2d(np.asarray(V))
    delta = np.pi / 180 * 7.5
    L = V.shape[0]
    t = np.linspace(0, 2 * np.pi, S)
    theta = 8 * t
    r = 1.05 ** theta
    r = (r - r.min()) / (r.max() - r.min())
    blocks = []
    for i in range(1, L + 1):
        cx = r * np.cos(theta + (i - 1) * delta)
        cy = r * np.sin(theta + (i - 1) * delta)
        cx = _round_half_away(cx * N / 2) + N / 2 + 1
        cy = _round_half_away(cy * N / 2) + N / 2 + 1
        cx = np.minimum(cx, N)
        cy = np.minimum(cy, N)
        ind = (cx + N * (cy - 1)).astype(np.int64)           # 1-based linear index into zeros(N)
        temp = np.zeros((N, N))
        temp.reshape(-1, order="F")[...] = 0
        tf = temp.reshape(-1, order="F").copy()
        tf[ind - 1] = 1
        temp = np.fft.fftshift(tf.reshape((N, N), order="F"))
        ind = np.flatnonzero(temp.reshape(-1, order="F") == 1) + 1   # find(temp==1), 1-based ascending
        n = ind.size
        tmp = sp.csr_matrix((np.ones(n), (np.arange(n), ind - 1)), shape=(n, N * M))
        blocks.append(tmp @ sp.kron(sp.csr_matrix(np.conj(V[i - 1:i, :])), sp.identity(N * M, format="csr"), format="csr"))
    return sp.vstack(blocks, format="csr")


def epi_P(N, M, percentage, V):
    """setup_subsampling_epi.m:20-35, line by line."""
    V = np.atleast_2d(np.asarray(V))
    step = int(_round_half_away(np.float64(1 / percentage)))
    no_of_steps = int(np.floor(N / step))
    nb_meas = no_of_steps * M
    L = V.shape[0]
    comb = np.zeros(N)
    comb[np.arange(1, step * nb_meas // M + 1, step) - 1] = 1        # comb(1:step:step*nb_meas/M) = 1
    blocks = []
    for _ in range(L):
        comb = comb[np.concatenate([[N], np.arange(1, N)]) - 1]     # comb([N,1:N-1])
        template = np.outer(comb, np.ones(M))
        ind = np.flatnonzero(template.reshape(-1, order="F") == 1) + 1
        n = ind.size
        tmp = sp.csr_matrix((np.ones(n), (np.arange(n), ind - 1)), shape=(n, N * M))
        i = len(blocks)
        blocks.append(tmp @ sp.kron(sp.csr_matrix(np.conj(V[i:i + 1, :])), sp.identity(N * M, format="csr"), format="csr"))
    return sp.vstack(blocks, format="csr")


class LiteralF:
    """main_recon_tsmis_FFT.m:228-229 with the explicit sparse P."""

    def __init__(self, P, N, M):
        self.P = P.tocsr()
        self.PH = P.conj().T.tocsr()
        self.N, self.M = N, M

    def forward(self, x):
        return self.P @ np.fft.fft2(x, axes=(0, 1)).reshape(-1, order="F") / np.sqrt(self.N * self.M)

    def adjoint(self, y):
        k = (self.PH @ y).reshape((self.N, self.M, -1), order="F")
        return np.fft.ifft2(k, axes=(0, 1)) * np.sqrt(self.N * self.M)


def lsqr_xupdate(F, y, v, uold, r, cg_tol=1e-4, maxit=100, x=None):
    """PnP_ADMM.m:102:  x = lsqr(@afun, [y(:); (v(:)-uold(:))*sqrt(r)], cg_tol, 100, [], [], x(:)).

    ``afun`` (:153-171): notransp  [F.forward(x); sqrt(r) x(:)],  transp  F.adjoint(y1) + sqrt(r) y2.
    scipy's lsqr takes two tolerances; MATLAB's single ``tol`` bounds the relative residual norm, so atol = btol = tol.
    """
    dims = v.shape
    s = np.asarray(y).size
    n = int(np.prod(dims))
    sr = np.sqrt(r)

    def mv(xv):
        X = xv.reshape(dims, order="F")
        return np.concatenate([F.forward(X).reshape(-1), xv * sr])

    def rmv(w):
        return F.adjoint(w[:s]).reshape(-1, order="F") + w[s:] * sr

    A = LinearOperator((s + n, n), matvec=mv, rmatvec=rmv, dtype=np.complex128)
    b = np.concatenate([np.asarray(y, np.complex128).reshape(-1),
                        (np.asarray(v, np.complex128).reshape(-1, order="F") - np.broadcast_to(np.asarray(uold, np.complex128), dims).reshape(-1, order="F")) * sr])
    x0 = None if x is None else np.asarray(x, np.complex128).reshape(-1, order="F")
    sol = lsqr(A, b, atol=cg_tol, btol=cg_tol, iter_lim=maxit, x0=x0)
    return sol[0].reshape(dims, order="F"), sol[2]
