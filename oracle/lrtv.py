"""Oracle: the LRTV comparison baseline (FISTA with a total-variation prox).

Test infrastructure (see ``oracle/__init__.py``).  NumPy float64 restatement of

* ``main_files/algorithms/LRTV/FISTA_deep.m:31-103`` - FISTA on ``0.5 |y - F.forward(x)|^2 + K |x|_TV`` with backtracking,
  the momentum ``(t-1)/(t+2)`` and the relative-objective stop of the reference; the TV term acts on the 2N x (M L) real image
  ``[reshape(real(x),N,[]); reshape(imag(x),N,[])]`` (real part stacked on the imaginary part, channels side by side - so the
  gradient also couples neighbouring channels and the real / imaginary seam, a quirk kept as is);
* the proximal operator it calls, ``unlocbox/prox/prox_tv.m:100-193`` (Beck & Teboulle's fast gradient projection on the dual,
  IEEE TIP 18(11) 2009), with the forward-difference ``gradient_op`` / its adjoint ``div_op`` / isotropic ``norm_tv`` of
  ``unlocbox/utils`` and the toolbox defaults the reference leaves in place (``tol = 10e-4``, ``maxit = 200``, unit weights,
  the toolbox's own momentum ``t = (1 + sqrt(4 t_old^2)) / 2``).  unlocbox is GPL third-party code vendored by the reference:
  the algorithm is restated from its published description and the call sites, nothing is copied.
"""
from __future__ import annotations

import numpy as np


def gradient_op(I):
    dx = np.zeros_like(I)
    dy = np.zeros_like(I)
    dx[:-1, :] = I[1:, :] - I[:-1, :]
    dy[:, :-1] = I[:, 1:] - I[:, :-1]
    return dx, dy


def div_op(dx, dy):
    I = np.empty_like(dx)
    I[0, :] = dx[0, :]
    I[1:-1, :] = dx[1:-1, :] - dx[:-2, :]
    I[-1, :] = -dx[-2, :]
    I[:, 0] += dy[:, 0]
    I[:, 1:-1] += dy[:, 1:-1] - dy[:, :-2]
    I[:, -1] += -dy[:, -2]
    return I


def norm_tv(I):
    dx, dy = gradient_op(I)
    return float(np.sum(np.sqrt(dx ** 2 + dy ** 2)))


def prox_tv(b, gamma, tol=10e-4, maxit=200):
    """argmin_z 0.5 |b - z|^2 + gamma |z|_TV; returns (sol, iterations)."""
    if gamma == 0:
        return b.copy(), 0
    r = np.zeros_like(b)
    s = np.zeros_like(b)
    pold, qold = r, s
    told, prev_obj = 1.0, 0.0
    it = 0
    for it in range(1, maxit + 1):
        sol = b - gamma * div_op(r, s)
        obj = 0.5 * np.sum((b - sol) ** 2) + gamma * norm_tv(sol)
        rel_obj = abs(obj - prev_obj) / obj
        prev_obj = obj
        if rel_obj < tol:
            break
        dx, dy = gradient_op(sol)
        r = r - dx / (8 * gamma)
        s = s - dy / (8 * gamma)
        w = np.maximum(1.0, np.sqrt(r ** 2 + s ** 2))
        p, q = r / w, s / w
        t = (1 + np.sqrt(4 * told ** 2)) / 2
        r = p + (told - 1) / t * (p - pold)
        s = q + (told - 1) / t * (q - qold)
        pold, qold, told = p, q, t
    return sol, it


def stack(x):
    """[reshape(real(x),N,[]); reshape(imag(x),N,[])] for x of shape N x M x L (column-major reshape)."""
    N = x.shape[0]
    return np.concatenate([np.real(x).reshape((N, -1), order="F"), np.imag(x).reshape((N, -1), order="F")], axis=0)


def unstack(x2, shape):
    N = shape[0]
    return (x2[:N, :] + 1j * x2[N:, :]).reshape(shape, order="F")


def fista_lrtv(y, F, shape, K=4e-5, max_iter=200, step=None, tol=1e-4, backtrack=True, trace=None):
    """``FISTA_deep(data, param)`` with ``data.{N,M,L,y,F}`` and ``param.{K,iter,step,tol,backtrack}``
    (``main_recon_tsmis_FFT.m:272-282``: K = 4e-5, iter = 200, step = numel(X0)/numel(Y), tol = 1e-4, backtrack = 1)."""
    y = np.asarray(y, np.complex128).reshape(-1)
    if step is None:
        step = float(np.prod(shape)) / y.size
    x = np.zeros(shape, np.complex128)
    t = 1
    x2_prev = x
    obj_prev = 0.0
    its = 0
    for its in range(1, max_iter + 1):
        err = F.forward(x).reshape(-1) - y
        grad1 = F.adjoint(err)
        cvxobj = 0.5 * np.linalg.norm(err) ** 2
        val = norm_tv(stack(x))
        while True:
            x2 = x - grad1 * step
            if K > 0:
                x2s, _ = prox_tv(stack(x2), step * K)
                x2 = unstack(x2s, shape)
            if not backtrack:
                break
            tmp = 0.5 * np.linalg.norm(F.forward(x2).reshape(-1) - y) ** 2
            d = (x2 - x).reshape(-1)
            if tmp > cvxobj + np.real(np.vdot(grad1.reshape(-1), d)) + np.linalg.norm(d) ** 2 / (2 * step):
                step = step / 2
            else:
                break
        x = x2 + (t - 1) / (t + 2) * (x2 - x2_prev)
        x2_prev = x2
        t += 1
        obj = cvxobj + K * val
        if trace is not None:
            trace.append({"obj": obj, "step": step})
        if abs(obj - obj_prev) / obj < tol:
            break
        obj_prev = obj
    return x, its
