"""Oracle / fixtures: synthetic dictionaries, QMaps, TSMI synthesis and noise.

Test infrastructure (see ``oracle/__init__.py``).  The reference ships neither
the dictionaries (``dictionaries/real_fisp_cut*_dict``) nor the ground-truth
QMaps, so every input is synthesised here, seeded.

* ``make_dictionary``      - FISP-like surrogate signals on a (T1,T2) grid,
  rank-C SVD compression -> ``dict.{D,normD,lut,V}`` with the field meanings of
  ``mrf_dtm_cpu.m:8-12`` / ``main_recon_tsmis_FFT.m:127-129``.
* ``make_qmaps``           - nested-ellipse brain phantom, ``[S x 3 x 230 x 230]``
  (layout of ``main_synthesize_tsmis.m:80``).
* ``synthesize_tsmis``     - restates ``main_synthesize_tsmis.m:84-98``
  (nearest (T1,T2) atom, scale by normD*|PD|, sign-align to channel 1).
* ``awgn_measured``        - restates ``awgn(Y, snr, 'measured')``
  (``main_recon_tsmis_FFT.m:243``) with a seeded generator.
"""
from __future__ import annotations

import numpy as np

CUT_T = {0: 1000, 1: 500, 2: 300, 3: 200, 4: 100}  # main_recon_tsmis_FFT.m:42


def _fisp_surrogate(T1, T2, T):
    """Cheap two-state FISP-like recursion (not a Bloch/EPG simulator).

    Inversion-prepared longitudinal state with sinusoidal flip-angle lobes and a
    partially refocused transverse memory term so that both T1 and T2 shape the
    fingerprint.  Only smoothness in (T1,T2) matters for the fixtures.
    """
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
    for i in range(T):
        a = fa[i]
        E1 = np.exp(-TR[i] / T1)
        E2 = np.exp(-TR[i] / T2)
        s = mz * np.sin(a) + mxy * np.cos(a / 2) ** 2
        sig[:, i:i + 1] = s * np.exp(-TE / T2)
        mz_new = mz * np.cos(a) - 0.5 * mxy * np.sin(a)
        mxy = 0.6 * s * E2
        mz = 1.0 - (1.0 - mz_new) * E1
    return sig


def make_dictionary(K_target=10000, cut=3, C=10, seed=0, t1_range=(0.1, 4.0), t2_range=(0.01, 0.6)):
    """Synthetic ``dict`` struct: D [K x C] unit-norm fp32, normD [K], lut [K x 2], V [T x C]."""
    T = CUT_T[cut]
    # log-spaced grid with T2 < T1; choose grid sizes to land near K_target
    n1 = int(np.ceil(np.sqrt(K_target * 1.6)))
    n2 = int(np.ceil(K_target * 1.3 / n1))
    t1 = np.geomspace(t1_range[0], t1_range[1], n1)
    t2 = np.geomspace(t2_range[0], t2_range[1], n2)
    g1, g2 = np.meshgrid(t1, t2, indexing="ij")
    keep = g2 < g1
    lut = np.stack([g1[keep], g2[keep]], axis=1)
    if lut.shape[0] > K_target:
        lut = lut[np.linspace(0, lut.shape[0] - 1, K_target).round().astype(int)]
    sig = _fisp_surrogate(lut[:, 0], lut[:, 1], T)
    # rank-C temporal subspace from a seeded subsample (cheap, deterministic)
    rng = np.random.default_rng(seed)
    sub = sig[rng.choice(sig.shape[0], size=min(2000, sig.shape[0]), replace=False)]
    _, _, vt = np.linalg.svd(sub, full_matrices=False)
    V = vt[:C].T                                        # T x C
    Dc = sig @ V                                        # K x C
    normD = np.linalg.norm(Dc, axis=1)
    D = Dc / normD[:, None]
    return {"D": D.astype(np.float32), "normD": normD.astype(np.float32),
            "lut": lut.astype(np.float32), "V": V.astype(np.float64)}


def make_qmaps(seed=0, S=15, N=230, M=230):
    """Nested-ellipse phantom: qmap [S x 3 x N x M] with rows (T1, T2, PD); zero background."""
    rng = np.random.default_rng(seed)
    yy, xx = np.meshgrid(np.linspace(-1, 1, N), np.linspace(-1, 1, M), indexing="ij")
    tissues = [  # (T1, T2, PD)
        (0.35, 0.07, 0.55),   # scalp/fat-like
        (1.40, 0.10, 0.85),   # grey matter
        (0.85, 0.07, 0.70),   # white matter
        (3.50, 0.50, 1.00),   # CSF
        (1.10, 0.09, 0.78),   # deep grey
    ]
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


def synthesize_tsmis(dict_, qmap_slice, C=10):
    """``main_synthesize_tsmis.m:84-98`` for one slice ``[3 x N x M]`` -> X ``[N x M x C]`` (real)."""
    from scipy.spatial import cKDTree  # Mdl = KDTreeSearcher(dict.lut), :54
    _, N, M = qmap_slice.shape
    qm = np.transpose(qmap_slice, (1, 2, 0)).reshape((-1, 3), order="F")   # pixel n + N*m
    tree = cKDTree(np.asarray(dict_["lut"], np.float64))
    _, I = tree.query(qm[:, :2].astype(np.float64), k=1)                   # knnsearch, :88
    X = np.real(np.asarray(dict_["D"], np.float64)[I, :C]) * np.asarray(dict_["normD"], np.float64)[I, None]
    X = X * np.abs(qm[:, 2:3])                                             # :93
    X = X.reshape((N, M, C), order="F")
    s = np.sign(X[:, :, 0:1])                                              # :97-98
    return X * s, I


def awgn_measured(Y, snr_db, seed):
    """``awgn(Y, snr, 'measured')``: complex white noise of power mean(|Y|^2)/10^(snr/10)."""
    rng = np.random.default_rng(seed)
    Y = np.asarray(Y)
    p = np.mean(np.abs(Y) ** 2) / (10.0 ** (snr_db / 10.0))
    n = (rng.standard_normal(Y.shape) + 1j * rng.standard_normal(Y.shape)) * np.sqrt(p / 2.0)
    return Y + n
