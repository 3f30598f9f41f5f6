"""Oracle: foreground mask and quality metrics of the driver script.

Test infrastructure (see ``oracle/__init__.py``).  Restates

* ``main_files/utils/getmask_fromPD.m:1-15`` - ``imfill(mask, 8, 'holes')`` of the thresholded grey image followed by
  ``mask(mask>0) = 1``: a pixel ends up 0 exactly when it is 0 after thresholding and the image border can reach it
  through 8-connected zero pixels (``scipy.ndimage.binary_fill_holes`` with the 3 x 3 structuring element floods the
  background the same way);
* ``main_recon_tsmis_FFT.m:328-384`` - masked MAE, ``psnr`` (peak value 1 for double images) and ``ssim`` with
  MATLAB's defaults: Gaussian window radius ``ceil(3*1.5) = 5`` (11 x 11), sigma 1.5, ``'replicate'`` padding,
  dynamic range 1, ``C = [(0.01)^2 (0.03)^2]``, the simplified (eq. 13) map averaged over all pixels.  MATLAB's
  ``psnr.m`` / ``ssim.m`` are toolbox sources absent from the reference checkout (Image Processing Toolbox, R2021a per
  ``README.md:71``): restated from their documentation - **parity unpinned** like the rest of the MATLAB half.
"""
from __future__ import annotations

import numpy as np
from scipy import ndimage


def getmask_fromPD(PD, thresh):
    pd = np.abs(np.asarray(PD)).astype(np.float64)
    pd = pd / pd.max()
    pd[pd < thresh] = 0
    filled = ndimage.binary_fill_holes(pd > 0, structure=np.ones((3, 3), bool))
    return filled.astype(np.float64)


def psnr(A, ref):
    err = np.mean((np.asarray(A, np.float64) - np.asarray(ref, np.float64)) ** 2)
    with np.errstate(divide="ignore"):
        return float(10.0 * np.log10(1.0 / err))


def ssim(A, ref):
    A = np.asarray(A, np.float64)
    ref = np.asarray(ref, np.float64)
    radius = 1.5
    fr = int(np.ceil(radius * 3))
    k = np.arange(-fr, fr + 1, dtype=np.float64)
    g = np.exp(-(k ** 2) / (2 * radius ** 2))
    g /= g.sum()

    def filt(X):
        return ndimage.correlate1d(ndimage.correlate1d(X, g, axis=0, mode="nearest"), g, axis=1, mode="nearest")

    C1, C2 = 0.01 ** 2, 0.03 ** 2
    mux, muy = filt(A), filt(ref)
    muxy, mux2, muy2 = mux * muy, mux ** 2, muy ** 2
    sx2 = filt(A * A) - mux2
    sy2 = filt(ref * ref) - muy2
    sxy = filt(A * ref) - muxy
    m = ((2 * muxy + C1) * (2 * sxy + C2)) / ((mux2 + muy2 + C1) * (sx2 + sy2 + C2))
    return float(m.mean())


def recon_metrics(qmap, qmap0, foreground_mask, X=None, X0=None):
    """main_recon_tsmis_FFT.m:328-384 -> dict under the script's variable names."""
    qmap = np.asarray(qmap).astype(np.complex128) if np.iscomplexobj(qmap) else np.asarray(qmap, np.float64)
    qmap0 = np.asarray(qmap0, np.float64)
    fm = np.ones(qmap.shape[:2]) if foreground_mask is None else np.asarray(foreground_mask, np.float64)
    ind = fm > 0
    t1 = np.real(qmap[:, :, 0] * fm)
    t1_ref = qmap0[:, :, 0] * fm
    t2 = np.real(qmap[:, :, 1] * fm)
    t2_ref = qmap0[:, :, 1] * fm
    pd = qmap[:, :, 2] * fm
    pd = np.abs(pd) / np.abs(pd).max()
    pd_ref = qmap0[:, :, 2] * fm
    pd_ref = np.abs(pd_ref) / np.abs(pd_ref).max()
    out = {}
    for nm, a, b in (("t1", t1, t1_ref), ("t2", t2, t2_ref), ("pd", pd, pd_ref)):
        out[nm + "_mae"] = float(np.mean(np.abs(a[ind] - b[ind])))
        out[nm + "_psnr"] = psnr(a, b)
        out[nm + "_ssim"] = ssim(a, b)
    if X is not None:
        X, X0 = np.asarray(X), np.asarray(X0)
        C = X0.shape[2]
        out["tsmi_mean_psnr"] = float(np.mean([psnr(np.abs(X[:, :, c]), np.abs(X0[:, :, c])) for c in range(C)]))
        out["tsmi_mean_ssim"] = float(np.mean([ssim(np.abs(X[:, :, c]), np.abs(X0[:, :, c])) for c in range(C)]))
    return out


# ---- measurement noise: the product's generator restated (Philox-4x32-10 + Box-Muller), see csrc/aux_kernels.cu ----
def philox4x32_10(ctr, key):
    """Philox-4x32-10 (Salmon, Moraes, Dror, Shaw: "Parallel random numbers: as easy as 1, 2, 3", SC'11).
    ``ctr``: uint32 array [..., 4]; ``key``: two uint32.  Known answer: ctr = key = 0 -> 6627e8d5 e169c58d bc57ac4c 9b00dbd8."""
    c = np.array(ctr, dtype=np.uint64, copy=True)
    k0, k1 = np.uint64(key[0]), np.uint64(key[1])
    M0, M1 = np.uint64(0xD2511F53), np.uint64(0xCD9E8D57)
    mask = np.uint64(0xFFFFFFFF)
    for _ in range(10):
        p0 = M0 * c[..., 0]
        p1 = M1 * c[..., 2]
        n0 = ((p1 >> np.uint64(32)) ^ c[..., 1] ^ k0) & mask
        n1 = p1 & mask
        n2 = ((p0 >> np.uint64(32)) ^ c[..., 3] ^ k1) & mask
        n3 = p0 & mask
        c = np.stack([n0, n1, n2, n3], axis=-1)
        k0 = (k0 + np.uint64(0x9E3779B9)) & mask
        k1 = (k1 + np.uint64(0xBB67AE85)) & mask
    return c.astype(np.uint32)


def awgn_philox(Y, snr_db, seed):
    """The noise ``qmri_awgn`` adds: per column s of Y (nmeas x S), sample i draws Philox(counter = (i, 0, s, 0), key = seed)."""
    Y = np.asarray(Y, np.complex128)
    Y2 = Y.reshape(Y.shape[0], -1)
    nmeas, S = Y2.shape
    out = np.empty_like(Y2)
    for s in range(S):
        ctr = np.zeros((nmeas, 4), np.uint64)
        ctr[:, 0] = np.arange(nmeas) & 0xFFFFFFFF
        ctr[:, 1] = np.arange(nmeas) >> 32
        ctr[:, 2] = s
        r = philox4x32_10(ctr, (seed & 0xFFFFFFFF, (seed >> 32) & 0xFFFFFFFF)).astype(np.float64)
        u1 = (r[:, 0] + 0.5) / 4294967296.0
        u2 = (r[:, 1] + 0.5) / 4294967296.0
        p = np.mean(np.abs(Y2[:, s]) ** 2) / 10.0 ** (snr_db / 10.0)
        rad = np.sqrt(-2.0 * np.log(u1)) * np.sqrt(p / 2.0)
        out[:, s] = Y2[:, s] + rad * (np.cos(2 * np.pi * u2) + 1j * np.sin(2 * np.pi * u2))
    return out.reshape(Y.shape)
