"""Seeded synthetic inputs for bench.py and the tests: the reference ships neither its dictionaries
(`dictionaries/real_fisp_cut*_dict`, .MISSING_LARGE_BLOBS) nor the ground-truth QMaps (GE data, README.md:85).

* ``make_dictionary``  FISP-like surrogate fingerprints on a log-spaced (T1, T2) grid with T2 < T1, rank-10 SVD compression ->
  ``dict.{D, normD, lut, V}`` with the field meanings of ``mrf_dtm_cpu.m:8-12`` / ``main_recon_tsmis_FFT.m:127-129``;
  ``cut`` selects the sequence length T of ``main_recon_tsmis_FFT.m:42`` ({1000, 500, 300, 200, 100} for cut0..cut4).
  ``atoms=(a0, a1)`` renders only that atom range (atom-sharded ranks build just their shard; the temporal basis V comes
  from a seeded subsample of the grid, so every rank derives the same V).
* ``make_qmaps``       nested-ellipse brain phantom, ``[S x 3 x 230 x 230]`` (layout of ``main_synthesize_tsmis.m:80``).
* ``volunteer_slices`` the 8 volunteers x 15 slices of the north-star job, cropped ``4:227`` (``main_recon_tsmis_FFT.m:187``).

Pure NumPy, no GPU, no oracle import: both benchmark arms and the tests draw their inputs from here.
"""
from __future__ import annotations

import numpy as np

CUT_T = {0: 1000, 1: 500, 2: 300, 3: 200, 4: 100}  # main_recon_tsmis_FFT.m:42


def fisp_surrogate(T1, T2, T):
    """Cheap two-state FISP-like recursion (not a Bloch / EPG simulator): inversion-prepared longitudinal state, sinusoidal
    flip-angle lobes, a partially refocused transverse memory term - both T1 and T2 shape the fingerprint smoothly."""
    t = np.arange(T)
    fa = np.deg2rad(10.0 + 50.0 * np.abs(np.sin(np.pi * t / 180.0)) * (0.6 + 0.4 * np.cos(np.pi * t / 410.0)))
    TR = 0.012 + 0.002 * np.sin(2 * np.pi * t / 97.0)
    TE = 0.004
    T1 = np.asarray(T1, np.float64)[:, None]
    T2 = np.asarray(T2, np.float64)[:, None]
    K = T1.shape[0]
    mz = -np.ones((K, 1))
    mxy = np.zeros((K, 1))
    sig = np.empty((K, T))
    e_te = np.exp(-TE / T2)
    for i in range(T):
        a = fa[i]
        E1 = np.exp(-TR[i] / T1)
        E2 = np.exp(-TR[i] / T2)
        s = mz * np.sin(a) + mxy * np.cos(a / 2) ** 2
        sig[:, i:i + 1] = s * e_te
        mz_new = mz * np.cos(a) - 0.5 * mxy * np.sin(a)
        mxy = 0.6 * s * E2
        mz = 1.0 - (1.0 - mz_new) * E1
    return sig


def dictionary_grid(K_target, t1_range=(0.1, 4.0), t2_range=(0.01, 0.6)):
    """The (T1, T2) look-up table: log-spaced grid, T2 < T1, thinned to K_target rows."""
    n1 = int(np.ceil(np.sqrt(K_target * 1.6)))
    n2 = int(np.ceil(K_target * 1.3 / n1))
    t1 = np.geomspace(t1_range[0], t1_range[1], n1)
    t2 = np.geomspace(t2_range[0], t2_range[1], n2)
    g1, g2 = np.meshgrid(t1, t2, indexing="ij")
    keep = g2 < g1
    lut = np.stack([g1[keep], g2[keep]], axis=1)
    if lut.shape[0] > K_target:
        lut = lut[np.linspace(0, lut.shape[0] - 1, K_target).round().astype(int)]
    return lut


def make_dictionary(K_target=10000, cut=3, C=10, seed=0, atoms=None, block=65536):
    """Synthetic ``dict`` struct: D [K x C] unit-norm fp32, normD [K], lut [K x 2], V [T x C].
    ``atoms=(a0, a1)``: D holds rows a0..a1-1 only (normD / lut stay whole; normD is NaN outside the range)."""
    T = CUT_T[cut]
    lut = dictionary_grid(K_target)
    K = lut.shape[0]
    rng = np.random.default_rng(seed)
    sub = np.sort(rng.choice(K, size=min(2000, K), replace=False))
    _, _, vt = np.linalg.svd(fisp_surrogate(lut[sub, 0], lut[sub, 1], T), full_matrices=False)
    V = vt[:C].T                                        # T x C temporal subspace
    a0, a1 = (0, K) if atoms is None else (int(atoms[0]), int(atoms[1]))
    D = np.empty((a1 - a0, C), np.float32)
    normD = np.full(K, np.nan, np.float32)
    for b0 in range(a0, a1, block):
        b1 = min(a1, b0 + block)
        Dc = fisp_surrogate(lut[b0:b1, 0], lut[b0:b1, 1], T) @ V
        nd = np.linalg.norm(Dc, axis=1)
        D[b0 - a0:b1 - a0] = (Dc / nd[:, None]).astype(np.float32)
        normD[b0:b1] = nd.astype(np.float32)
    return {"D": D, "normD": normD, "lut": lut.astype(np.float32), "V": V.astype(np.float64), "K": K, "atoms": (a0, a1)}


def make_qmaps(seed=0, S=15, N=230, M=230):
    """Nested-ellipse phantom: qmap [S x 3 x N x M] with rows (T1, T2, PD); zero background."""
    rng = np.random.default_rng(seed)
    yy, xx = np.meshgrid(np.linspace(-1, 1, N), np.linspace(-1, 1, M), indexing="ij")
    tissues = [(0.35, 0.07, 0.55), (1.40, 0.10, 0.85), (0.85, 0.07, 0.70), (3.50, 0.50, 1.00), (1.10, 0.09, 0.78)]
    q = np.zeros((S, 3, N, M), np.float64)
    for s in range(S):
        sc = 0.75 + 0.2 * np.sin(np.pi * (s + 0.5) / S)
        label = np.zeros((N, M), int)
        shapes = [(0.0, 0.0, 0.95 * sc, 0.80 * sc, 1), (0.0, 0.0, 0.88 * sc, 0.73 * sc, 2),
                  (0.02, 0.0, 0.70 * sc, 0.55 * sc, 3), (0.0, 0.0, 0.25 * sc, 0.12 * sc, 4),
                  (0.3 * sc, 0.25 * sc, 0.12, 0.10, 5), (0.3 * sc, -0.25 * sc, 0.12, 0.10, 5)]
        for (cy, cx, ry, rx, lab) in shapes:
            label[((yy - cy) / ry) ** 2 + ((xx - cx) / rx) ** 2 <= 1.0] = lab
        ph = rng.uniform(0, 2 * np.pi, 4)
        smooth = 1.0 + 0.05 * (np.sin(3 * xx + ph[0]) * np.cos(2 * yy + ph[1]) + 0.5 * np.sin(5 * yy + ph[2]) * np.cos(4 * xx + ph[3])) / 1.5
        for lab, (t1, t2, pdv) in enumerate(tissues, start=1):
            m = label == lab
            q[s, 0][m] = t1 * smooth[m]
            q[s, 1][m] = t2 * smooth[m]
            q[s, 2][m] = np.clip(pdv * smooth[m], 0, 1)
    return q


def volunteer_slices(first, last):
    """Ground-truth QMaps of global slices [first, last) of the 8 x 15 = 120-slice job, cropped to 224 x 224:
    slice g belongs to volunteer g // 15 (phantom seed = volunteer), its slice g % 15.  Returns [n x 3 x 224 x 224]."""
    out = []
    cache = {}
    for g in range(first, last):
        vol, sl = divmod(g, 15)
        if vol not in cache:
            cache[vol] = make_qmaps(seed=vol, S=15)
        out.append(cache[vol][sl][:, 3:227, 3:227])   # qmap0((4:227),(4:227),:)
    return np.stack(out) if out else np.zeros((0, 3, 224, 224))
