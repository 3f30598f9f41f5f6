"""The real-state PnP-ADMM loop (csrc/xupdate_real.cu; opt-in with QMRI_K1_STATE=real - it moves 8 instead of 24
bytes per pixel-channel but measured no faster than the complex-state kernels, which stay the default): the loop carries the
sampled-location recurrence c_{k+1} = (y - 2 m_k + m_{k-1} + c_k) / (1 + rho), m_k = A v_k, and only real images cross HBM; the
complex iterate x_K is materialised after the last iteration.  Checked here against the oracle loop (float64, exact solve) and
against the complex-state kernels (QMRI_K1_STATE=complex) for every iteration count that takes a different code path
(1: first x-update only; 2: base X0; 3: first + steady-free last; 4+: steady iterations), with a GENERAL complex X0 (not A^H y, so
the first x-update is a real solve and c_1 != 0), on both masks, and with slice groups that do not divide the batch."""
import numpy as np
import pytest

from conftest import rel_l2
from test_gpu_parity import box_denoiser, make_problem, smooth_tsmi, TOL_XUPDATE

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def q():
    import qmri_b200
    qmri_b200.Context.default()
    return qmri_b200


def _ops(q, kind):
    from oracle import sampling
    V = np.eye(10)
    if kind == "spiral":
        return q.setup_subsampling_spiralgrided(224, 224, 771, V), sampling.setup_subsampling_spiralgrided(224, 224, 771, V)
    return q.setup_subsampling_epi(224, 224, 1 / 65, V), sampling.setup_subsampling_epi(224, 224, 1 / 65, V)


@pytest.mark.parametrize("kind", ["spiral", "epi"])
@pytest.mark.parametrize("iters", [1, 2, 3, 4, 7])
def test_real_state_loop_general_x0(q, kind, iters, monkeypatch):
    from oracle.admm import pnp_admm
    S = 3
    P, Po = _ops(q, kind)
    Fo, Xgt, Y, X0 = make_problem(Po, 60 + iters, S=S)
    Y[:, 1] *= 2.5                                                    # per-slice dynamic range
    X0 = Fo.adjoint(Y) + 0.3 * smooth_tsmi(70 + iters, S=S, cplx=True)  # not A^H y: the first x-update moves x
    param = {"iter": iters, "gamma": 0.05, "denoiser_type": "single_level"}
    monkeypatch.setenv("QMRI_K1_STATE", "real")
    monkeypatch.setenv("QMRI_K1R_GROUP", "2")                         # groups of 2 + 1 slices
    x = q.PnP_ADMM(Y, dict(param, F=q.fft_operator(P), net=box_denoiser, X0=X0))
    for s in range(S):
        xo = pnp_admm(Y[:, s], dict(param, F=Fo, net=box_denoiser, X0=X0[..., s]), solver="exact")
        assert rel_l2(x[..., s], xo) <= TOL_XUPDATE, (kind, iters, s)
    monkeypatch.delenv("QMRI_K1R_GROUP")
    x1 = q.PnP_ADMM(Y, dict(param, F=q.fft_operator(P), net=box_denoiser, X0=X0))
    assert rel_l2(x1, x) <= 1e-6                                      # grouping only changes the launch partition
    monkeypatch.setenv("QMRI_K1_STATE", "complex")
    xc = q.PnP_ADMM(Y, dict(param, F=q.fft_operator(P), net=box_denoiser, X0=X0))
    assert rel_l2(x, xc) <= 3e-6


def test_real_state_loop_builtin_net_graph_and_session_restart(q, monkeypatch):
    """Built-in UNetRes, enough iterations for the CUDA-graph replay; a second run() on the same session restarts from X0."""
    from oracle import unetres
    from oracle.admm import pnp_admm
    P, Po = _ops(q, "spiral")
    Fo, Xgt, Y, X0 = make_problem(Po, 81, S=2)
    sd = unetres.make_state_dict(10, seed=0)
    net = q.UNetRes(sd, in_nc=10)
    param = {"iter": 8, "gamma": 0.05, "X0": X0, "denoiser_type": "single_level", "F": q.fft_operator(P), "net": net}
    monkeypatch.setenv("QMRI_K1_STATE", "real")
    sess = q.AdmmSession(param, 2)
    assert sess.xupdate_bytes() == 8
    sess.upload(Y, X0)
    sess.run(8)
    xa = sess.download()
    sess.run(3)
    sess.run(8)
    xb = sess.download()
    sess.close()
    assert np.array_equal(xa, xb)
    xo = pnp_admm(Y[:, 1], {"iter": 8, "gamma": 0.05, "F": Fo, "X0": X0[..., 1],
                            "net": lambda v: unetres.denoise_matlab_layout(sd, v)}, solver="exact")
    assert rel_l2(xa[..., 1], xo) <= 1e-4
    monkeypatch.setenv("QMRI_NO_GRAPH", "1")
    xd = q.PnP_ADMM(Y, param)
    assert np.array_equal(xa, xd)
    monkeypatch.setenv("QMRI_K1_STATE", "complex")
    xc = q.PnP_ADMM(Y, param)
    # two fp32 formulations of the same iteration, 8 passes through a random-init (non-contractive) network: the denoiser
    # tolerance bounds their distance, as it bounds each one's distance to the oracle
    assert rel_l2(xa, xc) <= 1e-4
    xo0 = pnp_admm(Y[:, 0], {"iter": 8, "gamma": 0.05, "F": Fo, "X0": X0[..., 0],
                             "net": lambda v: unetres.denoise_matlab_layout(sd, v)}, solver="exact")
    assert rel_l2(xc[..., 0], xo0) <= 1e-4 and rel_l2(xa[..., 0], xo0) <= 1e-4
