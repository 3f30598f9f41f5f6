"""The MEX gateway a MATLAB maintainer would build (mex/qmri_b200_mex.cpp) is compiled against a functional mock of the MEX API
- no MATLAB / Octave exists here - in BOTH complex storage models (interleaved = `mex -R2018a`, split = legacy MEX / Octave) and
EXECUTED: every command is driven through mexFunction and compared with the ctypes path (VERDICT r1 boundary item 6)."""
import os
import shutil
import subprocess

import numpy as np
import pytest

from conftest import PKG, ROOT, rel_l2


@pytest.mark.skipif(shutil.which("g++") is None, reason="g++ not available")
@pytest.mark.parametrize("flags", [[], ["-DQMRI_MOCK_SPLIT_COMPLEX"]])
def test_mex_gateway_compiles_against_header(flags):
    r = subprocess.run(["g++", "-std=c++17", "-fsyntax-only", "-Wall", "-DQMRI_MOCK_MEX", *flags, "-I" + os.path.join(PKG, "mex"),
                        "-I" + os.path.join(ROOT, "include"), os.path.join(PKG, "mex", "qmri_b200_mex.cpp")],
                       capture_output=True, text=True)
    assert r.returncode == 0, r.stderr


def test_matlab_wrappers_keep_reference_names():
    names = {"PnP_ADMM.m", "mrf_dtm_cpu.m", "setup_subsampling_spiralgrided.m", "setup_subsampling_epi.m",
             "denoiseImage_PnP_ADMM.m", "build_noise_map.m", "getmask_fromPD.m"}
    assert names <= set(os.listdir(os.path.join(PKG, "matlab")))
    for f in ("setup_subsampling_spiralgrided.m", "setup_subsampling_epi.m"):       # the constructors return the reference's handles
        txt = open(os.path.join(PKG, "matlab", f)).read()
        assert ".for = @(x)" in txt and ".adj = @(x)" in txt


@pytest.mark.parametrize("split", [False, True])
def test_mex_gateway_dispatch_and_error_mapping(split):
    """CPU-side: the gateway library loads, dispatches, and library errors surface as MATLAB exceptions (no GPU needed)."""
    from mexmock import MexError, MexMock
    m = MexMock(split)
    with pytest.raises(MexError) as e:
        m.call("no_such_command")
    assert e.value.ident == "qmri:usage"
    with pytest.raises(MexError) as e:
        m.call("forward")                                   # too few arguments
    assert e.value.ident == "qmri:usage"
    # the mock's own array round trip (what every GPU comparison below relies on)
    rng = np.random.default_rng(0)
    for a in (rng.standard_normal((3, 4, 2)), (rng.standard_normal((5, 2)) + 1j * rng.standard_normal((5, 2))).astype(np.complex64)):
        mx = m.to_mx(a)
        assert np.array_equal(m.from_mx(mx), a)
        m.lib.mxDestroyArray(mx)
    try:
        import torch
        has_gpu = torch.cuda.is_available()
    except Exception:
        has_gpu = False
    if not has_gpu:
        with pytest.raises(MexError) as e:
            m.call("op_spiral", 224, 224, 771, np.eye(10))
        assert e.value.ident == "qmri:error" and "no CPU fallback" in str(e.value)


@pytest.mark.gpu
@pytest.mark.parametrize("split", [False, True])
def test_mex_gateway_matches_ctypes_path(split):
    import benchdata
    import qmri_b200 as q
    from mexmock import FunctionHandle, MexError, MexMock
    m = MexMock(split)
    rng = np.random.default_rng(0)
    V = np.eye(10)
    h = m.call("op_spiral", 224, 224, 771, V)
    assert h.dtype == np.uint64 and m.lib.mock_is_locked()
    P = q.setup_subsampling_spiralgrided(224, 224, 771, V)
    F = q.fft_operator(P)
    assert int(m.call("nmeas", h)[0, 0]) == P.nmeas and np.array_equal(m.call("last_op"), h)
    X = rng.standard_normal((224, 224, 10, 2)) + 1j * rng.standard_normal((224, 224, 10, 2))
    y = m.call("forward", h, X)
    assert y.shape == (P.nmeas, 2) and np.array_equal(y, F.forward(X))
    assert np.array_equal(m.call("forward", h, X[..., 0].real)[:, 0], F.forward(X[..., 0].real))        # real input, one slice
    x = m.call("adjoint", h, y, np.array([224.0, 224.0, 10.0]))
    assert x.shape == (224, 224, 10, 2) and np.array_equal(x, F.adjoint(y))
    # K.for / K.adj and the reference's own closure lines (main_recon_tsmis_FFT.m:228-229) built from them
    kvec = np.fft.fft2(X[..., 0], axes=(0, 1)).reshape(-1, order="F")
    yf = m.call("p_for", h, kvec)[:, 0]
    assert np.array_equal(yf, P.for_(kvec))
    assert rel_l2(yf / 224.0, F.forward(X[..., 0])) < 1e-5
    va = m.call("p_adj", h, yf, 224 * 224 * 10)[:, 0]
    assert np.array_equal(va, P.adj(yf))
    # noise, mask, metrics
    yn = m.call("awgn", y, 30.0, 77)
    assert np.array_equal(yn, q.awgn(y, 30.0, "measured", seed=77))
    qm0 = np.transpose(benchdata.volunteer_slices(2, 3)[0], (1, 2, 0))
    mask = m.call("mask", qm0[:, :, 2], 0.15)
    assert np.array_equal(mask, q.getmask_fromPD(qm0[:, :, 2], 0.15))
    est = (qm0 * (1 + 0.05 * rng.standard_normal(qm0.shape))).astype(np.complex64)
    Xr = np.abs(X[..., 0]) * 0.1
    met = m.call("metrics", est, qm0, mask, Xr + 0.01, Xr)[:, 0]
    ref = q.recon_metrics(est, qm0, mask, Xr + 0.01, Xr)
    assert np.array_equal(met, np.array([ref[k] for k in q.metrics.METRIC_NAMES]))

    # the loop: function-handle denoiser (feval hop) and the built-in network handle
    def box(v):   # computed in double whatever the caller hands over (MATLAB feval gets a double array, the ctypes callback a single one)
        v = np.asarray(v, dtype=np.float64)
        return 0.5 * v + 0.125 * (np.roll(v, 1, 0) + np.roll(v, -1, 0) + np.roll(v, 1, 1) + np.roll(v, -1, 1))
    Y1 = yn[:, 0]
    X0 = F.adjoint(Y1)
    prm = {"iter": 4, "gamma": 0.05, "cg_tol": 1e-4, "X0": X0, "denoiser_type": "single_level"}
    xm = m.call("pnp_admm", h, Y1, dict(prm, net=FunctionHandle(box)))
    xq = q.PnP_ADMM(Y1, dict(prm, F=F, net=box))
    assert xm.shape == (224, 224, 10) and np.array_equal(xm, xq)
    from oracle import unetres
    sd = unetres.make_state_dict(10, seed=0)
    # a MATLAB caller hands every weight tensor over in PyTorch MEMORY order: the array [kw kh Cin Cout] (column-major) of the
    # tensor [Cout Cin kh kw] (row-major) - here the transposed view, which is Fortran-contiguous over the same bytes
    hn = m.call("net_load", 10, [sd[k].numpy().T for k, _ in q.state_dict_keys(10)])
    m.call("net_precision", hn, 1, nlhs=0)
    net = q.UNetRes(sd, in_nc=10)
    net.set_precision("tc")
    xm = m.call("pnp_admm", h, Y1, dict(prm, net=hn))
    assert np.array_equal(xm, q.PnP_ADMM(Y1, dict(prm, F=F, net=net)))
    A = rng.random((224, 224, 10))
    assert np.array_equal(m.call("denoise", hn, A), net.denoise(A))
    # matching
    d = benchdata.make_dictionary(K_target=3000, cut=3, seed=0)
    hd = m.call("dict_load", d["D"], d["normD"], d["lut"])
    xs = np.asfortranarray(xq.reshape((-1, 10), order="F")[:4096].astype(np.complex64))
    qmap, pd, mt, dm = m.call("match", hd, xs, 2, nlhs=4)
    out = q.mrf_dtm_cpu(d, {"X": xs.reshape((4096, 1, 10))}, {"f": {"qout": 1, "pdout": 1, "mtout": 1, "dmout": 1}})
    assert np.array_equal(qmap, out["qmap"][:, 0, :]) and np.array_equal(pd[:, 0], out["pd"][:, 0])
    assert np.array_equal(mt[:, 0], out["mt"][:, 0]) and np.array_equal(dm[:, 0], out["dm"][:, 0].astype(np.int32))
    # errors from the library become MATLAB exceptions
    with pytest.raises(MexError) as e:
        m.call("op_spiral", 128, 128, 771, V)
    assert e.value.ident == "qmri:error" and "224" in str(e.value)
    m.call("shutdown", nlhs=0)
