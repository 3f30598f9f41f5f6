"""Oracle vs golden vectors (CPU).  Pins the NumPy restatement against SURVEY.md Appendix A and
against self-consistency identities - the only pins available for the MATLAB parts."""
import json
import os

import numpy as np
import pytest

from conftest import GOLDEN, rel_l2
from oracle import sampling
from oracle.admm import pnp_admm, pnp_admm_wv
from oracle.matching import mrf_dtm_cpu
from oracle.xupdate import xupdate_closed_form, xupdate_exact, xupdate_lsqr

A = json.load(open(os.path.join(GOLDEN, "appendix_a.json")))


def X0_a2():
    n = np.arange(1, 225)[:, None, None]
    m = np.arange(1, 225)[None, :, None]
    c = np.arange(1, 11)[None, None, :]
    return np.sin(0.1 * n + 0.05 * c) * np.cos(0.07 * m) + 0.01 * c


def test_spiral_mask_known_answers():
    fr = sampling.spiral_frame_indices(224, 224, 771, 10)
    assert [len(f) for f in fr] == A["spiral_counts"]
    assert [int((f + 1).sum()) for f in fr] == A["spiral_idx_sums_1based"]
    assert list(fr[0][:5] + 1) == A["spiral_first5_frame1"] and list(fr[0][-3:] + 1) == A["spiral_last3_frame1"]
    assert list(fr[9][:5] + 1) == A["spiral_first5_frame10"] and list(fr[9][-3:] + 1) == A["spiral_last3_frame10"]
    assert len(np.unique(np.concatenate(fr))) == 2686
    assert all(0 in f for f in fr)  # DC always sampled
    fr200 = sampling.spiral_frame_indices(224, 224, 771, 200)
    cnt = [len(f) for f in fr200]
    assert (min(cnt), max(cnt), sum(cnt)) == (613, 621, 123604)
    assert np.array_equal(fr200[48], fr200[0])  # 48-frame rotation period (7.5 degrees)


def test_epi_mask_known_answers():
    fr = sampling.epi_frame_indices(224, 224, 1 / 65, 10)
    assert all(len(f) == A["epi_count"] for f in fr)
    for k, v in A["epi_idx_sums_1based"].items():
        assert int((fr[int(k) - 1] + 1).sum()) == v
    assert list(fr[0][:5] + 1) == A["epi_first5_frame1"]
    rows = sorted(set(int(k % 224) for k in fr[3]))
    assert rows == [4, 69, 134]  # rows {1+i, 66+i, 131+i}, 1-based, for frame i = 4


def test_matlab_round_half_away_from_zero():
    assert list(sampling.matlab_round([0.5, 1.5, 2.5, -0.5, -1.5])) == [1, 2, 3, -1, -2]


@pytest.mark.parametrize("name", ["spiral", "epi"])
def test_operator_and_xupdate_known_answers(name):
    g = A[name]
    X0 = X0_a2()
    assert abs((X0 ** 2).sum() - A["X0_norm2"]) < 1e-6 * A["X0_norm2"]
    P = (sampling.setup_subsampling_spiralgrided(224, 224, 771, np.eye(10)) if name == "spiral"
         else sampling.setup_subsampling_epi(224, 224, 1 / 65, np.eye(10)))
    F = sampling.FOperator(P)
    y = F.forward(X0)
    assert P.nmeas == g["nmeas"]
    cx = lambda p: p[0] + 1j * p[1]
    assert abs((abs(y) ** 2).sum() - g["y_norm2"]) < 1e-11 * g["y_norm2"]
    for got, key in ((y[0], "y1"), (y[1], "y2"), (y[-1], "yend")):
        assert abs(got - cx(g[key])) < 1e-11
    x0 = F.adjoint(y)
    assert abs(x0[0, 0, 0] - cx(g["x0_111"])) < 1e-11 and abs(x0[223, 223, 9] - cx(g["x0_end"])) < 1e-11
    assert abs((abs(x0) ** 2).sum() - g["y_norm2"]) < 1e-9 * g["y_norm2"]  # A A^H = I
    z = 0.9 * X0 + 0.1j * np.roll(X0, 3, axis=0)
    x = xupdate_exact(F, y, z, 0.05)
    assert abs((abs(x) ** 2).sum() - g["x_norm2"]) < 1e-11 * g["x_norm2"]
    assert abs(x[0, 0, 0] - cx(g["x_111"])) < 1e-11 and abs(x[99, 49, 4] - cx(g["x_100_50_5"])) < 1e-11
    # three formulations agree
    assert rel_l2(xupdate_closed_form(F, y, z, 0.05), x) < 1e-14
    xl, its = xupdate_lsqr(F, y, z, 0.05, tol=1e-4, x0=x0)
    assert its == 2 and rel_l2(xl, x) < 1e-13


def test_adjoint_identity_and_measurement_order():
    rng = np.random.default_rng(0)
    V = np.linalg.qr(rng.standard_normal((10, 10)))[0]  # orthonormal rows, channel mixing
    P = sampling.setup_subsampling_spiralgrided(224, 224, 771, V)
    F = sampling.FOperator(P)
    a = rng.standard_normal((224, 224, 10)) + 1j * rng.standard_normal((224, 224, 10))
    b = rng.standard_normal(P.nmeas) + 1j * rng.standard_normal(P.nmeas)
    assert abs(np.vdot(b, F.forward(a)) - np.vdot(F.adjoint(b), a)) < 1e-9
    assert rel_l2(F.forward(F.adjoint(b)), b) < 1e-12          # A A^H = I for orthonormal rows
    assert P.rows_orthonormal()
    # frame-major, ascending k inside a frame
    for i in range(10):
        seg = P.idx[P.frame_ptr[i]:P.frame_ptr[i + 1]]
        assert np.all(np.diff(seg) > 0)


def test_general_V_exact_vs_lsqr():
    """T x 10 V with orthonormal columns: LSQR(1e-4) stops early, 7e-5 from the exact solve (A.5)."""
    rng = np.random.default_rng(1)
    V = np.linalg.qr(rng.standard_normal((24, 4)))[0]          # 24 frames, 4 channels (small for speed)
    P = sampling.setup_subsampling_spiralgrided(224, 224, 771, V)
    F = sampling.FOperator(P)
    assert not P.rows_orthonormal()
    xg = rng.standard_normal((224, 224, 4))
    y = F.forward(xg)
    z = 0.5 * xg
    xe = xupdate_exact(F, y, z, 0.05)
    # exact solve satisfies the normal equations
    r = F.adjoint(y - F.forward(xe)) - 0.05 * (xe - z)
    assert np.linalg.norm(r) < 1e-9 * np.linalg.norm(F.adjoint(y))
    xl, its = xupdate_lsqr(F, y, z, 0.05, tol=1e-12, maxit=200, x0=F.adjoint(y))
    assert rel_l2(xl, xe) < 1e-8


def test_matching_known_answers():
    g = A["match"]
    X0 = X0_a2()
    z = 0.9 * X0 + 0.1j * np.roll(X0, 3, axis=0)
    K = g["K"]
    k = np.arange(1, K + 1)[:, None]
    cc = np.arange(1, 11)[None, :]
    Draw = np.cos(0.37 * k * cc + 0.11 * cc ** 2) + 0.5 * np.sin(0.013 * k + cc)
    normD = np.linalg.norm(Draw, axis=1)
    lut = np.stack([0.1 + 3.9 * (k[:, 0] - 1) / (K - 1), 0.01 + 0.59 * np.mod(7 * (k[:, 0] - 1), K) / (K - 1)], 1)
    d = {"D": Draw / normD[:, None], "normD": normD, "lut": lut}
    o = mrf_dtm_cpu(d, {"X": z}, {"fp": {"blockSize": 1e9}}, return_gap=True)
    dm = o["dm"].reshape(-1, order="F")
    assert int(dm.sum()) == g["dm_sum"] and list(dm[:5]) == g["dm_first5"] and dm[24999] == g["dm_25000"]
    assert abs(o["qmap"][..., 0].astype(np.float64).sum() - g["T1_sum"]) < 1e-6 * g["T1_sum"]
    assert abs(o["qmap"][..., 1].astype(np.float64).sum() - g["T2_sum"]) < 1e-6 * g["T2_sum"]
    assert abs(np.abs(o["pd"]).astype(np.float64).sum() - g["abs_pd_sum"]) < 1e-6 * g["abs_pd_sum"]
    assert abs(o["pd"].reshape(-1, order="F")[0] - (g["pd1"][0] + 1j * g["pd1"][1])) < 1e-6
    gap = o["gap"].reshape(-1)
    assert int((gap < 1e-6).sum()) == g["n_gap_lt_1e-6"] and int((gap < 1e-4).sum()) == g["n_gap_lt_1e-4"]
    # block-size independence and the float32 mirror
    o2 = mrf_dtm_cpu(d, {"X": z}, {"fp": {"blockSize": 3e6}})
    assert np.array_equal(o2["dm"], o["dm"])
    o32 = mrf_dtm_cpu(d, {"X": z}, None, precision="f32")
    assert np.array_equal(o32["dm"], o["dm"])


def test_matching_ties_nan_and_4d():
    rng = np.random.default_rng(3)
    D = rng.standard_normal((32, 6))
    D[20] = D[3]
    D /= np.linalg.norm(D, axis=1, keepdims=True)
    lut = np.stack([np.arange(32.0), np.arange(32.0)], 1)
    lut[3, 1] = np.nan
    d = {"D": D, "normD": np.full(32, 2.0), "lut": lut}
    X = np.zeros((2, 3, 2, 6))
    X[...] = 4.0 * D[3]
    o = mrf_dtm_cpu(d, {"X": X})
    assert o["dm"].shape == (2, 3, 2) and np.all(o["dm"] == 4)
    assert np.all(o["qmap"][..., 1] == 0) and np.all(o["qmap"][..., 0] == 3)
    assert np.allclose(o["pd"], 2.0)


def test_state_reformulation_equals_faithful_loop():
    """(w, v) formulation == faithful (x, v, u) loop on the full-size operator."""
    from oracle.synth import awgn_measured
    P = sampling.setup_subsampling_epi(224, 224, 1 / 65, np.eye(3))
    F = sampling.FOperator(P)
    rng = np.random.default_rng(5)
    Xgt = rng.standard_normal((224, 224, 3))
    Y = awgn_measured(F.forward(Xgt), 30, 1)

    def net(v):
        return (v + np.roll(v, 1, 0) + np.roll(v, -1, 0) + np.roll(v, 1, 1) + np.roll(v, -1, 1)) / 5.0
    param = {"iter": 5, "gamma": 0.05, "F": F, "X0": F.adjoint(Y), "net": net}
    xa = pnp_admm(Y, param, solver="exact")
    xb = pnp_admm_wv(Y, param)
    xc = pnp_admm(Y, param, solver="lsqr")
    assert rel_l2(xb, xa) < 1e-13 and rel_l2(xc, xa) < 1e-12
