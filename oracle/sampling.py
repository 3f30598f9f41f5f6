"""Oracle: k-space subsampling patterns and the masked unitary FFT operator.

Test infrastructure (see ``oracle/__init__.py``).  NumPy float64 restatement of

* ``main_files/subsampling_patterns/setup_subsampling_spiralgrided.m:7-42``
* ``main_files/subsampling_patterns/setup_subsampling_epi.m:20-35``
* ``main_recon_tsmis_FFT.m:228-229`` (``F.forward`` / ``F.adjoint``)

All index arrays returned here are 0-based, ascending column-major linear
indices ``n + N*m`` (MATLAB's ``find`` order minus one).
"""
from __future__ import annotations

import numpy as np


def matlab_round(x):
    """MATLAB ``round``: halves away from zero (NumPy rounds halves to even)."""
    x = np.asarray(x, dtype=np.float64)
    return np.sign(x) * np.floor(np.abs(x) + 0.5)


def spiral_frame_indices(N: int, M: int, S: int, L: int):
    """Per-frame sampled locations of the gridded rotating spiral.

    ``setup_subsampling_spiralgrided.m:7-34``: log spiral r = 1.05^theta with
    theta = 8 t, t = linspace(0, 2 pi, S), rotated 7.5 degrees per frame,
    gridded with ``round``, clipped to N, de-duplicated through a 0/1 image,
    ``fftshift``-ed and listed with ``find``.  The reference builds an N x N
    image (``zeros(N)``, line 31) so N == M is assumed there; we mirror that.
    """
    if N != M:
        raise ValueError("setup_subsampling_spiralgrided assumes N == M (zeros(N), line 31)")
    delta = np.pi / 180.0 * 7.5
    t = np.linspace(0.0, 2.0 * np.pi, S)
    theta = 8.0 * t
    r = 1.05 ** theta
    r = (r - r.min()) / (r.max() - r.min())
    frames = []
    for i in range(L):
        cx = r * np.cos(theta + i * delta)
        cy = r * np.sin(theta + i * delta)
        cx = matlab_round(cx * N / 2) + N / 2 + 1
        cy = matlab_round(cy * N / 2) + N / 2 + 1
        cx = np.minimum(cx, N)
        cy = np.minimum(cy, N)
        ind = (cx + N * (cy - 1)).astype(np.int64) - 1  # 0-based linear, column-major
        temp = np.zeros(N * N, dtype=np.uint8)
        temp[ind] = 1
        temp = temp.reshape((N, N), order="F")
        temp = np.fft.fftshift(temp)
        frames.append(np.flatnonzero(temp.reshape(-1, order="F")).astype(np.int64))
    return frames


def epi_frame_indices(N: int, M: int, percentage: float, L: int):
    """Per-frame sampled locations of the multi-shot EPI comb.

    ``setup_subsampling_epi.m:20-30``: ``step = round(1/percentage)`` lines
    apart, ``floor(N/step)`` full lines, comb shifted by one row *before* the
    first frame is taken, applied to unshifted k-space.
    """
    step = int(matlab_round(1.0 / percentage))
    no_of_steps = N // step
    comb = np.zeros(N, dtype=np.uint8)
    comb[0:step * no_of_steps:step] = 1  # comb(1:step:step*nb_meas/M) = 1
    frames = []
    for _ in range(L):
        comb = np.roll(comb, 1)  # comb([N,1:N-1])
        template = np.repeat(comb[:, None], M, axis=1)  # comb*ones(1,M)
        frames.append(np.flatnonzero(template.reshape(-1, order="F")).astype(np.int64))
    return frames


class SubsamplingOperator:
    """The sparse matrix ``P`` of the reference, applied matrix-free.

    Row block i of P is ``S_i * kron(conj(V(i,:)), I_NM)``
    (``setup_subsampling_spiralgrided.m:36-37``, ``setup_subsampling_epi.m:31-32``):
    frame i samples ``sum_c conj(V[i,c]) * xhat_c`` on its own mask.  Measurement
    order: frame-major, ascending column-major k index inside a frame.
    """

    def __init__(self, N, M, frames, V):
        self.N, self.M = int(N), int(M)
        self.V = np.atleast_2d(np.asarray(V))
        self.L, self.C = self.V.shape
        if len(frames) != self.L:
            raise ValueError("need one index list per row of V")
        self.frames = [np.asarray(f, dtype=np.int64) for f in frames]
        self.frame_ptr = np.concatenate([[0], np.cumsum([len(f) for f in self.frames])]).astype(np.int64)
        self.idx = np.concatenate(self.frames) if self.L else np.zeros(0, np.int64)
        self.nmeas = int(self.frame_ptr[-1])

    # P * x   (x is the column-major vec of an N x M x C array)
    def for_(self, x):
        x = np.asarray(x).reshape(-1)
        xc = x.reshape(self.C, self.N * self.M)  # channel c at offset c*N*M
        out = np.empty(self.nmeas, dtype=np.result_type(x.dtype, self.V.dtype, np.complex128))
        for i, f in enumerate(self.frames):
            out[self.frame_ptr[i]:self.frame_ptr[i + 1]] = np.conj(self.V[i]) @ xc[:, f]
        return out

    # P' * y  (conjugate transpose)
    def adj(self, y):
        y = np.asarray(y).reshape(-1)
        out = np.zeros((self.C, self.N * self.M), dtype=np.result_type(y.dtype, self.V.dtype, np.complex128))
        for i, f in enumerate(self.frames):
            out[:, f] += self.V[i][:, None] * y[self.frame_ptr[i]:self.frame_ptr[i + 1]][None, :]
        return out.reshape(-1)

    def rows_orthonormal(self, tol=1e-10):
        G = self.V @ self.V.conj().T
        return bool(np.max(np.abs(G - np.eye(self.L))) < tol)


def setup_subsampling_spiralgrided(N, M, S, V):
    V = np.atleast_2d(np.asarray(V))
    return SubsamplingOperator(N, M, spiral_frame_indices(N, M, S, V.shape[0]), V)


def setup_subsampling_epi(N, M, percentage, V):
    V = np.atleast_2d(np.asarray(V))
    return SubsamplingOperator(N, M, epi_frame_indices(N, M, percentage, V.shape[0]), V)


class FOperator:
    """``F.forward`` / ``F.adjoint`` of ``main_recon_tsmis_FFT.m:228-229``.

    Arrays are handled in the MATLAB layout N x M x C (fft2 over the first two
    dims).  A leading batch of slices is supported with an extra trailing dim
    N x M x C x S, which equals S independent reference calls.
    """

    def __init__(self, P: SubsamplingOperator):
        self.P = P
        self.N, self.M, self.C = P.N, P.M, P.C

    def forward(self, x):
        x = np.asarray(x)
        if x.ndim == 4:
            return np.stack([self.forward(x[..., s]) for s in range(x.shape[3])], axis=1)
        xh = np.fft.fft2(x.reshape(self.N, self.M, self.C), axes=(0, 1))
        return self.P.for_(xh.reshape(-1, order="F")) / np.sqrt(self.N * self.M)

    def adjoint(self, y):
        y = np.asarray(y)
        if y.ndim == 2:
            return np.stack([self.adjoint(y[:, s]) for s in range(y.shape[1])], axis=3)
        k = self.P.adj(y).reshape((self.N, self.M, self.C), order="F")
        return np.fft.ifft2(k, axes=(0, 1)) * np.sqrt(self.N * self.M)
