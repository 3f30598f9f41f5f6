"""Oracle: the least-squares (x-update) step of PnP-ADMM.

Test infrastructure (see ``oracle/__init__.py``).  Restates ``PnP_ADMM.m:102``

    x = lsqr(@afun, [y(:); (v(:)-uold(:))*sqrt(r)], cg_tol, 100, [], [], x(:))

with ``afun`` from ``PnP_ADMM.m:153-171``, i.e. the minimiser of
``||y - A x||^2 + rho ||x - z||^2`` with ``z = v - uold``, ``A = F.forward``.

Three formulations are provided:

* ``xupdate_exact``        - normal equations solved exactly in k-space; the
  C x C block ``G_k + rho I`` per sampled location (any V).  This is the truth
  the CUDA path is compared with (truncated LSQR(1e-4) is
  7e-5 away from it for general V, so it cannot be the 1e-5 yardstick).
* ``xupdate_closed_form``  - ``z + A^H (y - A z)/(1+rho)``, valid when
  ``A A^H = I`` (rows of V orthonormal).
* ``xupdate_lsqr``         - Paige-Saunders LSQR on the stacked operator with
  a warm start, the reference's actual numerical route.  MATLAB's ``lsqr.m`` is
  proprietary and absent; the stopping rule used here (relative residual OR
  relative normal-equation residual <= tol) is a restatement from memory and is
  only used to show agreement with the exact solve.
"""
from __future__ import annotations

import numpy as np

from .sampling import FOperator


def xupdate_closed_form(F: FOperator, y, z, rho):
    z = np.asarray(z, dtype=np.complex128)
    return z + F.adjoint(np.asarray(y) - F.forward(z)) / (1.0 + rho)


def xupdate_exact(F: FOperator, y, z, rho):
    P = F.P
    N, M, C = F.N, F.M, F.C
    z = np.asarray(z, dtype=np.complex128).reshape(N, M, C)
    zh = np.fft.fft2(z, axes=(0, 1)) / np.sqrt(N * M)
    rhs = P.adj(np.asarray(y, dtype=np.complex128)).reshape(C, N * M) + rho * zh.reshape(-1, order="F").reshape(C, N * M)
    xh = rhs / rho
    # normal-matrix blocks on the union of sampled locations
    G = {}
    for i, f in enumerate(P.frames):
        blk = np.outer(P.V[i], np.conj(P.V[i]))  # V(i,:)^T conj(V(i,:))
        for k in f:
            G[k] = G.get(k, 0) + blk
    eye = np.eye(C)
    for k, g in G.items():
        xh[:, k] = np.linalg.solve(g + rho * eye, rhs[:, k])
    xh = xh.reshape(-1).reshape((N, M, C), order="F")
    return np.fft.ifft2(xh, axes=(0, 1)) * np.sqrt(N * M)


def xupdate_lsqr(F: FOperator, y, z, rho, tol=1e-4, maxit=100, x0=None):
    """Complex Paige-Saunders LSQR on B = [A; sqrt(rho) I], rhs [y; sqrt(rho) z]."""
    N, M, C = F.N, F.M, F.C
    sr = np.sqrt(rho)
    y = np.asarray(y, dtype=np.complex128).reshape(-1)
    z = np.asarray(z, dtype=np.complex128).reshape(N, M, C)
    s = y.size

    def Aop(x):  # afun 'notransp'
        x = x.reshape((N, M, C), order="F")
        return np.concatenate([F.forward(x).reshape(-1), x.reshape(-1, order="F") * sr])

    def ATop(w):  # afun 'transp'
        return F.adjoint(w[:s]).reshape(-1, order="F") + w[s:] * sr

    b = np.concatenate([y, z.reshape(-1, order="F") * sr])
    x = np.zeros(N * M * C, np.complex128) if x0 is None else np.asarray(x0, np.complex128).reshape(-1, order="F").copy()
    bnorm = np.linalg.norm(b)
    u = b - Aop(x)
    beta = np.linalg.norm(u)
    its = 0
    if beta == 0 or beta / bnorm <= tol:
        return x.reshape((N, M, C), order="F"), its
    u = u / beta
    v = ATop(u)
    alpha = np.linalg.norm(v)
    if alpha == 0:
        return x.reshape((N, M, C), order="F"), its
    v = v / alpha
    w = v.copy()
    phibar, rhobar = beta, alpha
    anorm2 = 0.0
    for its in range(1, maxit + 1):
        u = Aop(v) - alpha * u
        beta = np.linalg.norm(u)
        if beta > 0:
            u = u / beta
        anorm2 += alpha ** 2 + beta ** 2
        v = ATop(u) - beta * v
        alpha = np.linalg.norm(v)
        if alpha > 0:
            v = v / alpha
        rho_ = np.hypot(rhobar, beta)
        c, sn = rhobar / rho_, beta / rho_
        theta = sn * alpha
        rhobar = -c * alpha
        phi = c * phibar
        phibar = sn * phibar
        x = x + (phi / rho_) * w
        w = v - (theta / rho_) * w
        rnorm = phibar
        arnorm = phibar * alpha * abs(c)
        if rnorm / bnorm <= tol or arnorm / (np.sqrt(anorm2) * max(rnorm, 1e-300)) <= tol:
            break
    return x.reshape((N, M, C), order="F"), its


def norm_zero_to_one(x):
    """``PnP_ADMM.m:174-184``: global min/max over the whole array, no zero-range guard."""
    x_min = x.min()
    x_max = x.max()
    x_range = x_max - x_min
    return (x - x_min) / x_range, x_min, x_max, x_range


def undo_norm_zero_to_one(x, x_min, x_max, x_range):
    """``PnP_ADMM.m:187-192``."""
    return x * x_range + x_min
