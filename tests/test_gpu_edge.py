"""Edge cases of the C ABI on the GPU: empty inputs (zero pixels, zero slices), one pixel, and the error behaviour the header promises -
negative status + message, never a crash, never a silently different computation (include/qmri.h; INTEGRATION.md "Limits")."""
import ctypes as C

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def q():
    import qmri_b200
    qmri_b200.Context.default()
    return qmri_b200


def test_empty_and_single_pixel_matching(q):
    import benchdata
    from oracle.matching import mrf_dtm_cpu as oracle_match
    d = benchdata.make_dictionary(K_target=3000, cut=3, seed=5)
    par = {"f": {"qout": 1, "pdout": 1, "dmout": 1, "mtout": 1}}
    out = q.mrf_dtm_cpu(d, {"X": np.zeros((0, 1, 10), np.complex64)}, par)          # no pixels: empty outputs, no launch failure
    assert out["qmap"].shape == (0, 1, 2) and out["pd"].shape == (0, 1) and out["dm"].shape == (0, 1)
    x1 = (d["D"][1234] * (0.7 - 0.2j)).reshape(1, 1, 10)                           # one pixel, exactly on an atom
    o1 = q.mrf_dtm_cpu(d, {"X": x1}, par)
    ref = oracle_match(d, {"X": x1}, None)
    assert int(o1["dm"][0, 0]) == int(ref["dm"][0, 0]) == 1235
    assert np.allclose(o1["qmap"], ref["qmap"]) and np.allclose(o1["pd"], ref["pd"], rtol=1e-5)
    with pytest.raises(ValueError):
        q.mrf_dtm_cpu(d, {"X": np.zeros((4, 4, 7))}, par)                           # timepoints != columns of dict.D


def test_zero_slices_and_bad_arguments(q):
    ctx = q.Context.default()
    lib = ctx.lib
    V = np.eye(10)
    P = q.setup_subsampling_spiralgrided(224, 224, 771, V)
    F = q.fft_operator(P)
    x = np.zeros((224, 224, 10, 1), np.complex128, order="F")
    y = np.zeros((P.nmeas, 1), np.complex128, order="F")
    vp = C.c_void_p
    # S = 0 is a no-op that succeeds (qmri.h)
    assert lib.qmri_forward(P.handle, vp(x.ctypes.data), q.QMRI_C128, 0, vp(y.ctypes.data), q.QMRI_C128) == 0
    assert lib.qmri_adjoint(P.handle, vp(y.ctypes.data), q.QMRI_C128, 0, vp(x.ctypes.data), q.QMRI_C128) == 0
    # null pointers and real-valued measurements are errors with a message, not crashes
    assert lib.qmri_forward(P.handle, None, q.QMRI_C128, 1, vp(y.ctypes.data), q.QMRI_C128) < 0
    assert lib.qmri_adjoint(P.handle, vp(y.ctypes.data), q.QMRI_F64, 1, vp(x.ctypes.data), q.QMRI_C128) < 0
    assert b"complex" in lib.qmri_last_error() or b"must" in lib.qmri_last_error()
    # shapes the wrappers refuse before the ABI is reached
    with pytest.raises(ValueError):
        F.forward(np.zeros((224, 224, 9)))
    with pytest.raises(ValueError):
        F.adjoint(np.zeros(P.nmeas + 1, np.complex128))


def test_unsupported_sizes_fail_loudly(q):
    with pytest.raises(q.QmriError, match="224"):
        q.setup_subsampling_spiralgrided(128, 128, 771, np.eye(10))                 # this build: N = M = 224 only
    with pytest.raises(q.QmriError, match="224"):
        q.setup_subsampling_epi(224, 192, 1 / 65, np.eye(10))
    P = q.setup_subsampling_spiralgrided(224, 224, 771, np.eye(4))                  # a 4-channel operator is fine ...
    F = q.fft_operator(P)
    Y = np.zeros(P.nmeas, np.complex128)
    with pytest.raises(q.QmriError, match="10"):                                    # ... but the PnP-ADMM loop needs the denoiser's 10 channels
        q.PnP_ADMM(Y, {"iter": 2, "gamma": 0.05, "F": F, "X0": F.adjoint(Y), "net": lambda v: v})


def test_admm_parameter_validation(q):
    P = q.setup_subsampling_spiralgrided(224, 224, 771, np.eye(10))
    F = q.fft_operator(P)
    Y = np.zeros(P.nmeas, np.complex128)
    X0 = F.adjoint(Y)
    with pytest.raises(q.QmriError, match="gamma"):
        q.PnP_ADMM(Y, {"iter": 2, "gamma": 0.0, "F": F, "X0": X0, "net": lambda v: v})
    with pytest.raises(KeyError, match="noise_map"):                                 # PnP_ADMM.m:70-72
        q.PnP_ADMM(Y, {"iter": 2, "gamma": 0.05, "F": F, "X0": X0, "net": lambda v: v[:, :, :10], "denoiser_type": "multi_level"})
    # an all-zero problem is well defined up to the 0/0 of norm_zero_to_one (PnP_ADMM.m:182 has no guard): zero iterations return X0
    x = q.PnP_ADMM(Y, {"iter": 0, "gamma": 0.05, "F": F, "X0": X0, "net": lambda v: v})
    assert x.shape == (224, 224, 10) and not np.any(x)


def test_operator_from_explicit_indices_rejects_bad_masks(q):
    ctx = q.Context.default()
    lib = ctx.lib
    h = C.c_void_p()
    idx = np.array([5, 3, 9], np.int32)                                             # not ascending inside the frame
    fp = np.array([0, 3], np.int64)
    rc = lib.qmri_op_create(ctx.handle, 224, 224, 1, 1, idx.ctypes.data_as(C.c_void_p), fp.ctypes.data_as(C.c_void_p), None, C.byref(h))
    assert rc < 0 and b"ascend" in lib.qmri_last_error()
