"""GPU parity tests: the CUDA path (through the C ABI) against the CPU oracle.

Tolerances (BASELINE.json north_star):
  * operator / x-update : relative L2 <= 1e-5 vs the float64 oracle
  * matching            : atom indices identical except oracle near-ties (top-2 relative gap < 1e-6)
  * denoiser            : relative L2 <= 1e-4 vs the fp32 CPU forward with the same weights
"""
import numpy as np
import pytest

from conftest import rel_l2

pytestmark = pytest.mark.gpu

TOL_XUPDATE = 1e-5
TOL_DENOISER = 1e-4
NEAR_TIE = 1e-6


@pytest.fixture(scope="module")
def q():
    import qmri_b200
    qmri_b200.Context.default()  # raises without a B200: no fallback
    return qmri_b200


@pytest.fixture(scope="module")
def ops(q):
    from oracle import sampling
    V = np.eye(10)
    out = {}
    out["spiral"] = (q.setup_subsampling_spiralgrided(224, 224, 771, V), sampling.setup_subsampling_spiralgrided(224, 224, 771, V))
    out["epi"] = (q.setup_subsampling_epi(224, 224, 1 / 65, V), sampling.setup_subsampling_epi(224, 224, 1 / 65, V))
    return out


def smooth_tsmi(seed, S=None, cplx=False):
    """Smooth, brain-like-ish random TSMI (N x M x C [x S])."""
    rng = np.random.default_rng(seed)
    shape = (224, 224, 10) + (() if S is None else (S,))
    n = np.arange(224)[:, None, None] / 224.0
    m = np.arange(224)[None, :, None] / 224.0
    c = np.arange(10)[None, None, :]
    base = np.exp(-((n - 0.5) ** 2 + (m - 0.45) ** 2) * 8) * np.cos(0.3 * c) + 0.2 * np.sin(7 * n + c) * np.cos(5 * m)
    x = base if S is None else np.repeat(base[..., None], S, axis=3)
    x = x + 0.05 * rng.standard_normal(shape)
    if cplx:
        x = x + 1j * (0.3 * np.roll(x, 5, axis=0) + 0.05 * rng.standard_normal(shape))
    return x


# ---------------------------------------------------------------------------------------------
@pytest.mark.parametrize("name", ["spiral", "epi"])
def test_mask_constructors_match_oracle(ops, name):
    P, Po = ops[name]
    idx, fp = P.indices()
    assert P.nmeas == Po.nmeas
    assert np.array_equal(fp, Po.frame_ptr)
    assert np.array_equal(idx, Po.idx)


@pytest.mark.parametrize("name", ["spiral", "epi"])
@pytest.mark.parametrize("dtype", [np.float64, np.complex128, np.float32, np.complex64])
def test_forward_adjoint(q, ops, name, dtype):
    from oracle.sampling import FOperator
    P, Po = ops[name]
    F, Fo = q.fft_operator(P), FOperator(Po)
    x = smooth_tsmi(1, cplx=np.issubdtype(dtype, np.complexfloating)).astype(dtype)
    y = F.forward(x)
    yo = Fo.forward(x.astype(np.complex128))
    assert y.shape == yo.shape
    assert rel_l2(y, yo) <= TOL_XUPDATE
    xa = F.adjoint(yo)
    xo = Fo.adjoint(yo)
    assert rel_l2(xa, xo) <= TOL_XUPDATE


def test_forward_adjoint_batched_and_adjointness(q, ops):
    from oracle.sampling import FOperator
    P, Po = ops["spiral"]
    F, Fo = q.fft_operator(P), FOperator(Po)
    x = smooth_tsmi(2, S=3, cplx=True)
    y = F.forward(x)
    assert y.shape == (P.nmeas, 3)
    for s in range(3):
        assert rel_l2(y[:, s], Fo.forward(x[..., s])) <= TOL_XUPDATE
    rng = np.random.default_rng(0)
    b = rng.standard_normal((P.nmeas, 3)) + 1j * rng.standard_normal((P.nmeas, 3))
    xa = F.adjoint(b)
    # <A x, b> == <x, A^H b>, and A A^H = I when V = eye
    lhs, rhs = np.vdot(b, y), np.vdot(xa, x)
    assert abs(lhs - rhs) <= 1e-5 * abs(lhs)
    assert rel_l2(F.forward(xa), b) <= TOL_XUPDATE


@pytest.mark.parametrize("name", ["spiral", "epi"])
def test_xupdate_matches_exact_solve(q, ops, name):
    from oracle.sampling import FOperator
    from oracle.xupdate import xupdate_exact
    P, Po = ops[name]
    F, Fo = q.fft_operator(P), FOperator(Po)
    x_true = smooth_tsmi(3)
    y = Fo.forward(x_true)
    v = smooth_tsmi(4)
    u = 0.1 * smooth_tsmi(5, cplx=True)
    rho = 0.05
    x, w, mm = F.xupdate(y, v, u, rho, want_w=True)
    xo = xupdate_exact(Fo, y, v - u, rho)
    assert rel_l2(x, xo) <= TOL_XUPDATE
    wo = xo + u
    assert rel_l2(w, wo) <= TOL_XUPDATE
    assert abs(mm[0] - wo.real.min()) <= 1e-5 * abs(wo.real).max()
    assert abs(mm[1] - wo.real.max()) <= 1e-5 * abs(wo.real).max()
    # u omitted = the scalar 0 of PnP_ADMM.m:78
    x2 = F.xupdate(y, v, None, rho)
    assert rel_l2(x2, xupdate_exact(Fo, y, v, rho)) <= TOL_XUPDATE


def test_xupdate_appendix_a2_known_answers(q, ops):
    """SURVEY.md Appendix A.2 (independent float64 restatement): deterministic inputs."""
    n = np.arange(1, 225)[:, None, None]
    m = np.arange(1, 225)[None, :, None]
    c = np.arange(1, 11)[None, None, :]
    X0 = np.sin(0.1 * n + 0.05 * c) * np.cos(0.07 * m) + 0.01 * c
    z = 0.9 * X0 + 0.1j * np.roll(X0, 3, axis=0)
    expect = {"spiral": (3.124627321012e+04, 1.089099617982e+05, 6.803444208467e-01 + 2.255400393693e-02j),
              "epi": (6.178596834661e+03, 1.041382994863e+05, 6.691885559181e-01 + 6.113009617138e-02j)}
    for name, (ynorm2, xnorm2, x100_50_5) in expect.items():
        F = q.fft_operator(ops[name][0])
        y = F.forward(X0)
        assert abs((np.abs(y) ** 2).sum() - ynorm2) <= 2e-5 * ynorm2
        x = F.xupdate(y, z, None, 0.05)
        assert abs((np.abs(x) ** 2).sum() - xnorm2) <= 2e-5 * xnorm2
        assert abs(x[99, 49, 4] - x100_50_5) <= 1e-5 * abs(x100_50_5) * 10


# ---------------------------------------------------------------------------------------------
def a3_dictionary():
    K = 1000
    k = np.arange(1, K + 1)[:, None]
    cc = np.arange(1, 11)[None, :]
    Draw = np.cos(0.37 * k * cc + 0.11 * cc ** 2) + 0.5 * np.sin(0.013 * k + cc)
    normD = np.linalg.norm(Draw, axis=1)
    lut = np.stack([0.1 + 3.9 * (k[:, 0] - 1) / (K - 1), 0.01 + 0.59 * np.mod(7 * (k[:, 0] - 1), K) / (K - 1)], 1)
    return {"D": Draw / normD[:, None], "normD": normD, "lut": lut}


def check_match(out, ref, K):
    dm = out["dm"].reshape(-1, order="F").astype(np.int64)
    dmo = ref["dm"].reshape(-1, order="F")
    gap = ref["gap"].reshape(-1, order="F")
    decided = gap >= NEAR_TIE
    assert np.array_equal(dm[decided], dmo[decided]), f"{(dm[decided] != dmo[decided]).sum()} index mismatches outside near-ties"
    assert dm.min() >= 1 and dm.max() <= K
    same = dm == dmo
    q_, qo = out["qmap"].reshape(-1, out["qmap"].shape[-1], order="F"), ref["qmap"].reshape(-1, ref["qmap"].shape[-1], order="F")
    assert np.array_equal(q_[same], qo[same])
    pd, pdo = out["pd"].reshape(-1, order="F"), ref["pd"].reshape(-1, order="F")
    assert rel_l2(pd[same], pdo[same]) <= 1e-5
    return int((~same).sum())


def test_match_appendix_a3(q):
    """K = 1000 analytic dictionary, complex data, 22 near-tie pixels."""
    from oracle.matching import mrf_dtm_cpu as oracle_match
    n = np.arange(1, 225)[:, None, None]
    m = np.arange(1, 225)[None, :, None]
    c = np.arange(1, 11)[None, None, :]
    X0 = np.sin(0.1 * n + 0.05 * c) * np.cos(0.07 * m) + 0.01 * c
    z = 0.9 * X0 + 0.1j * np.roll(X0, 3, axis=0)
    d = a3_dictionary()
    par = {"f": {"qout": 1, "pdout": 1, "mtout": 1, "dmout": 1, "Xout": 0, "Yout": 0, "verbose": 0}, "fp": {"blockSize": 1e9}}
    out = q.mrf_dtm_cpu(d, {"X": z}, par)
    ref = oracle_match(d, {"X": z}, par, return_gap=True)
    assert out["qmap"].shape == (224, 224, 2) and out["pd"].shape == (224, 224)
    assert out["pd"].dtype == np.complex64 and out["qmap"].dtype == np.float32
    nmis = check_match(out, ref, 1000)
    assert nmis <= 22
    assert rel_l2(out["mt"], ref["mt"]) <= 1e-5


@pytest.mark.parametrize("cplx", [False, True])
def test_match_synthetic_dictionary(q, cplx):
    from oracle import synth
    from oracle.matching import mrf_dtm_cpu as oracle_match
    d = synth.make_dictionary(K_target=6000, cut=3, seed=1)
    K = d["D"].shape[0]
    rng = np.random.default_rng(7)
    # pixels = scaled noisy atoms (realistic: near-collinear neighbours compete) + a zero pixel + ragged count
    pick = rng.integers(0, K, size=3001)
    X = d["D"][pick].astype(np.float64) * rng.uniform(0.2, 2.0, size=(3001, 1)) + 0.01 * rng.standard_normal((3001, 10))
    if cplx:
        X = X * np.exp(1j * rng.uniform(0, 2 * np.pi, size=(3001, 1)))
    X[17] = 0
    X = X.reshape(3001, 1, 10)
    par = {"f": {"qout": 1, "pdout": 1, "mtout": 0, "dmout": 1}}
    out = q.mrf_dtm_cpu(d, {"X": X}, par)
    ref = oracle_match(d, {"X": X}, par, return_gap=True)
    check_match(out, ref, K)
    assert out["dm"].reshape(-1)[17] == 1  # all-zero pixel: MATLAB max returns the first index


def test_match_lut_nan_becomes_zero_and_first_index_ties(q):
    rng = np.random.default_rng(3)
    D = rng.standard_normal((64, 10))
    D[40] = D[5]      # exact duplicate atoms: the lower index must win
    D /= np.linalg.norm(D, axis=1, keepdims=True)
    lut = np.stack([np.arange(64, dtype=np.float64), np.arange(64, dtype=np.float64) * 2], 1)
    lut[5, 1] = np.nan
    d = {"D": D, "normD": np.ones(64), "lut": lut}
    X = (3.0 * D[5])[None, None, :]
    out = q.mrf_dtm_cpu(d, {"X": X}, {"f": {"qout": 1, "pdout": 1, "dmout": 1}})
    assert out["dm"][0, 0] == 6
    assert out["qmap"][0, 0, 0] == 5.0 and out["qmap"][0, 0, 1] == 0.0
    assert abs(out["pd"][0, 0] - 3.0) < 1e-5


# ---------------------------------------------------------------------------------------------
@pytest.fixture(scope="module")
def nets(q):
    from oracle import unetres
    sd = unetres.make_state_dict(10, seed=0)
    return q.UNetRes(sd, in_nc=10), sd


@pytest.mark.parametrize("precision", ["fp32", "tc"])
@pytest.mark.parametrize("shape", [(1, 32, 32), (2, 64, 40), (1, 224, 224), (9, 16, 24), (3, 56, 120)])
def test_denoiser_matches_cpu_forward(q, nets, shape, precision):
    """fp32 = CUDA-core exact mode; tc = tcgen05 split-bf16 tensor mode.  Same 1e-4 bar for both."""
    import torch
    from oracle import unetres
    net, sd = nets
    S, H, W = shape
    torch.manual_seed(11)
    x = torch.rand(S, 10, H, W)
    with torch.no_grad():
        ref = unetres.unetres_forward(sd, x).numpy()
    net.set_precision(precision)
    try:
        out = net.forward(x.numpy())
    finally:
        net.set_precision("tc")  # the library default
    assert out.shape == ref.shape
    err = rel_l2(out, ref)
    print(f"denoiser {precision} {shape}: rel-L2 {err:.3e}")
    assert err <= TOL_DENOISER


def test_denoiser_split_k_fallback_and_launch_modes(q, nets, monkeypatch):
    """One slice splits K at the 256- and 512-channel levels: through a thread-block cluster (default) or the L2 workspace with
    tickets (QMRI_NO_CLUSTER_SPLITK); with and without programmatic dependent launch.  All within the denoiser tolerance."""
    import torch
    from oracle import unetres
    net, sd = nets
    net.set_precision("tc")
    rng = np.random.default_rng(11)
    x = rng.random((1, 10, 224, 224)).astype(np.float32)
    with torch.no_grad():
        ref = unetres.unetres_forward(sd, torch.from_numpy(x)).numpy()
    y0 = net.forward(x)
    monkeypatch.setenv("QMRI_NO_CLUSTER_SPLITK", "1")
    y1 = net.forward(x)
    monkeypatch.setenv("QMRI_NO_PDL", "1")
    y2 = net.forward(x)
    monkeypatch.delenv("QMRI_NO_CLUSTER_SPLITK")
    y3 = net.forward(x)
    assert np.array_equal(y1, y2) and np.array_equal(y0, y3)      # dependent launch changes timing only
    assert rel_l2(y0, y1) <= 1e-5                                  # other split counts: another summation order
    assert rel_l2(y0, ref) <= TOL_DENOISER and rel_l2(y1, ref) <= TOL_DENOISER


def test_denoiser_matlab_layout_and_wrapper(q, nets):
    from oracle import unetres
    net, sd = nets
    rng = np.random.default_rng(5)
    A = rng.random((48, 32, 10))           # H x W x C, double like MATLAB
    ref = unetres.denoise_matlab_layout(sd, A)
    out = q.denoiseImage_PnP_ADMM(A, net, True, False)
    assert out.dtype == np.float64 and out.shape == (48, 32, 10)
    assert rel_l2(out, ref) <= TOL_DENOISER
    res = q.denoiseImage_PnP_ADMM(A, net, True, True)
    assert rel_l2(res, A - ref) <= 1e-4
    assert q.denoiseImage_PnP_ADMM(A, net, "yes", False) is None  # the reference's soft failure


def test_denoiser_chunked_passes_equal_single_pass(q, monkeypatch):
    """Slice batches beyond unetres.h's max_chunk run as balanced passes (QMRI_NET_CHUNK lowers the limit): 5 slices as 2 + 2 + 1
    against one pass of 5."""
    from oracle import unetres
    sd = unetres.make_state_dict(10, seed=0)
    rng = np.random.default_rng(5)
    A = rng.random((224, 224, 10, 5))
    one = q.UNetRes(sd, in_nc=10)
    y1 = one.denoise(A)
    monkeypatch.setenv("QMRI_NET_CHUNK", "2")
    three = q.UNetRes(sd, in_nc=10)
    y3 = three.denoise(A)
    assert y1.shape == y3.shape == (224, 224, 10, 5)
    assert rel_l2(y3, y1) <= 1e-5            # tile shapes / split-K follow the pass size: not bitwise
    ref = unetres.denoise_matlab_layout(sd, A[..., 4])
    assert rel_l2(y3[..., 4], ref) <= TOL_DENOISER


def test_denoiser_operand_swapped_64_channel_kernel(q, monkeypatch):
    """QMRI_TC_SWAP=1: the 64 -> 64 convs run through the operand-swapped kernel (weights resident in TMEM as the MMA's A operand,
    224 pixels as its N; csrc/conv_tc.cu conv64_swap) - same result as the CTA-pair kernel, within the denoiser tolerance of the CPU forward."""
    from oracle import unetres
    sd = unetres.make_state_dict(10, seed=0)
    rng = np.random.default_rng(6)
    A = rng.random((224, 224, 10, 3))
    net = q.UNetRes(sd, in_nc=10)
    net.set_precision("tc")
    ctx = q.Context.default()
    y_pair = net.denoise(A)
    monkeypatch.setenv("QMRI_TC_SWAP", "1")
    l0 = ctx.launch_count
    y_swap = net.denoise(A)
    assert ctx.launch_count - l0 >= 64
    assert rel_l2(y_swap, y_pair) <= 1e-5    # measured 4.4e-6: other summation order, the a_lo w_lo term on top (tensor mode itself: 5e-6 from the CPU forward)
    assert rel_l2(y_swap[..., 1], unetres.denoise_matlab_layout(sd, A[..., 1])) <= TOL_DENOISER


def test_denoiser_ab_switches_of_the_64_channel_path(q, monkeypatch):
    """The 64 -> 64 layers' data path (staging tile + bulk tensor stores / loads, one or two staging buffers) against its fallbacks:
    QMRI_TC_TMAOUT=0 (16-byte global accesses), QMRI_TC_OUTBUF=1 / 2 (forced buffer count) - the arithmetic and its order are the same,
    so the outputs are bit-identical - and QMRI_TC_HEADTAIL=0 (fp32 CUDA-core head / tail instead of pack + tensor conv + unpack)."""
    from oracle import unetres
    sd = unetres.make_state_dict(10, seed=0)
    rng = np.random.default_rng(8)
    A = rng.random((224, 224, 10, 2))
    net = q.UNetRes(sd, in_nc=10)
    net.set_precision("tc")
    y0 = net.denoise(A)
    for key, val in [("QMRI_TC_TMAOUT", "0"), ("QMRI_TC_OUTBUF", "1"), ("QMRI_TC_OUTBUF", "2")]:
        monkeypatch.setenv(key, val)
        y = net.denoise(A)
        monkeypatch.delenv(key)
        assert np.array_equal(y, y0), (key, val, rel_l2(y, y0))
    monkeypatch.setenv("QMRI_TC_HEADTAIL", "0")
    y_fp32_ht = net.denoise(A)
    monkeypatch.delenv("QMRI_TC_HEADTAIL")
    assert rel_l2(y_fp32_ht, y0) <= 1e-5
    assert rel_l2(y0[..., 1], unetres.denoise_matlab_layout(sd, A[..., 1])) <= TOL_DENOISER


def test_denoiser_multi_level_11_channels(q):
    import torch
    from oracle import unetres
    sd = unetres.make_state_dict(11, seed=0)
    net = q.UNetRes(sd, in_nc=11)
    torch.manual_seed(3)
    x = torch.rand(1, 11, 32, 32)
    x[:, 10] = 0.01
    with torch.no_grad():
        ref = unetres.unetres_forward(sd, x).numpy()
    assert rel_l2(net.forward(x.numpy()), ref) <= TOL_DENOISER


def test_synthesize_tsmis_matches_oracle(q):
    """main_synthesize_tsmis.m:84-98 on the GPU (exhaustive nearest-(T1,T2) scan) vs the cKDTree restatement.
    Indices must agree except where the float64 oracle's two nearest atoms are (nearly) equidistant; X follows from the index."""
    from scipy.spatial import cKDTree
    from oracle import synth
    d = synth.make_dictionary(K_target=9000, cut=3, seed=2)
    qmap = synth.make_qmaps(seed=3, S=2, N=230, M=230)[1]          # [3 x 230 x 230], zero background
    Xo, Io = synth.synthesize_tsmis(d, qmap)
    X, I = q.synthesize_tsmis(d, qmap, return_index=True)
    assert X.shape == (230, 230, 10) and X.dtype == np.float32
    qm = np.transpose(qmap, (1, 2, 0)).reshape((-1, 3), order="F")
    dist, _ = cKDTree(np.asarray(d["lut"], np.float64)).query(qm[:, :2].astype(np.float64), k=2)
    decided = (dist[:, 1] - dist[:, 0]) > 1e-5 * np.maximum(dist[:, 1], 1e-12)      # the GPU scan works on the float32 qmap
    assert decided.mean() > 0.3                                     # the phantom has large decided regions besides the background
    assert np.array_equal(I[decided], Io[decided])
    same = (I == Io).reshape((230, 230), order="F")
    assert rel_l2(X[same], Xo[same]) <= 1e-6
    assert np.all(X[:, :, 0] >= 0)                                  # channel 1 sign-aligned (:97-98)
    # the synthesized slice drives the rest of the path: crop 4:227 like main_recon_tsmis_FFT.m:212 and match it back
    out = q.mrf_dtm_cpu(d, {"X": X[3:227, 3:227, :]}, {"f": {"qout": 1, "pdout": 1, "dmout": 1}})
    fg = (np.abs(qmap[2, 3:227, 3:227]) > 0) & same[3:227, 3:227]
    lut = np.asarray(d["lut"])
    assert np.allclose(out["qmap"][fg][:, 0], lut[I.reshape((230, 230), order="F")[3:227, 3:227][fg], 0], rtol=0.03)


def test_match_atom_sharded_equals_unsharded(q):
    """BASELINE config 5 on one GPU: two atom shards scored separately, keys max-combined, finish -> identical to the
    unsharded match (the same packed keys are what ranks all-reduce with MAX; tests/test_sharding_gloo.py covers the exchange)."""
    import torch
    from oracle import synth
    d = synth.make_dictionary(K_target=6000, cut=3, seed=1)
    K = np.asarray(d["D"]).shape[0]
    rng = np.random.default_rng(4)
    npix = 5000
    xr = torch.from_numpy(rng.standard_normal((10, npix)).astype(np.float32)).cuda()
    xi = torch.from_numpy(rng.standard_normal((10, npix)).astype(np.float32)).cuda()
    full = q.Dictionary(d)
    qm0, pd0, mt0, dm0 = q.mrf_dtm_sharded(full, xr, xi, npix, want_mt=True)
    keys = []
    import ctypes as C
    for r in range(2):
        sh = q.Dictionary(d, shard=q.atom_shard(K, 2, r))
        k = torch.zeros(npix, dtype=torch.int64, device="cuda")
        q._capi.check(sh.ctx.lib.qmri_match_keys_dev(sh.handle, C.c_void_p(xr.data_ptr()), C.c_void_p(xi.data_ptr()), npix, C.c_void_p(k.data_ptr())))
        sh.ctx.synchronize()
        keys.append(k)
        sh.close()
    merged = torch.maximum(keys[0], keys[1])
    dm = torch.empty(npix, dtype=torch.int32, device="cuda")
    qm = torch.empty(2 * npix, dtype=torch.float32, device="cuda")
    pd = torch.empty(2 * npix, dtype=torch.float32, device="cuda")
    q._capi.check(full.ctx.lib.qmri_match_finish_dev(full.handle, C.c_void_p(xr.data_ptr()), C.c_void_p(xi.data_ptr()), npix,
                                                    C.c_void_p(merged.data_ptr()), C.c_void_p(qm.data_ptr()), C.c_void_p(pd.data_ptr()), None,
                                                    C.c_void_p(dm.data_ptr())))
    full.ctx.synchronize()
    assert torch.equal(dm, dm0) and torch.equal(qm.view(2, npix), qm0) and torch.equal(pd.view(npix, 2), pd0)
    score, idx = q.unpack_keys(merged.cpu().numpy().view(np.uint64))
    assert np.array_equal(idx + 1, dm0.cpu().numpy())
    full.close()


# ---------------------------------------------------------------------------------------------
def box_denoiser(v):
    """A cheap deterministic stand-in for param.net: 5-point average on the first 10 channels."""
    v = np.asarray(v, dtype=np.float64)[:, :, :10]
    return (v + np.roll(v, 1, 0) + np.roll(v, -1, 0) + np.roll(v, 1, 1) + np.roll(v, -1, 1)) / 5.0


def make_problem(Po, seed, S=None):
    from oracle.sampling import FOperator
    from oracle.synth import awgn_measured
    Fo = FOperator(Po)
    Xgt = smooth_tsmi(seed, S=S)
    Y = awgn_measured(Fo.forward(Xgt), 30, seed)
    return Fo, Xgt, Y, Fo.adjoint(Y)


@pytest.mark.parametrize("name", ["spiral", "epi"])
def test_admm_loop_with_callback_denoiser(q, ops, name):
    from oracle.admm import pnp_admm
    P, Po = ops[name]
    Fo, Xgt, Y, X0 = make_problem(Po, 21)
    param = {"iter": 6, "gamma": 0.05, "cg_tol": 1e-4, "gt_tsmi": Xgt, "X0": X0, "denoiser_type": "single_level"}
    xo = pnp_admm(Y, dict(param, F=Fo, net=box_denoiser), solver="exact")
    x = q.PnP_ADMM(Y, dict(param, F=q.fft_operator(P), net=box_denoiser))
    assert x.shape == (224, 224, 10) and x.dtype == np.complex128
    assert rel_l2(x, xo) <= TOL_XUPDATE


def test_admm_loop_batched_slices_are_independent(q, ops):
    from oracle.admm import pnp_admm
    P, Po = ops["spiral"]
    Fo, Xgt, Y, X0 = make_problem(Po, 22, S=2)
    Y[:, 1] *= 3.0  # different dynamic range per slice: min/max normalisation must be per slice
    X0 = Fo.adjoint(Y)
    param = {"iter": 4, "gamma": 0.05, "denoiser_type": "single_level"}
    x = q.PnP_ADMM(Y, dict(param, F=q.fft_operator(P), net=box_denoiser, X0=X0))
    for s in range(2):
        xo = pnp_admm(Y[:, s], dict(param, F=Fo, net=box_denoiser, X0=X0[..., s]), solver="exact")
        assert rel_l2(x[..., s], xo) <= TOL_XUPDATE


def test_admm_loop_batched_epi(q, ops):
    """BASELINE configs[2] shape in small: a slice batch on the EPI mask takes the streaming kernels (row chunks + overflow slots)."""
    from oracle.admm import pnp_admm
    P, Po = ops["epi"]
    Fo, Xgt, Y, X0 = make_problem(Po, 26, S=3)
    param = {"iter": 4, "gamma": 0.05, "denoiser_type": "single_level"}
    x = q.PnP_ADMM(Y, dict(param, F=q.fft_operator(P), net=box_denoiser, X0=X0))
    for s in range(3):
        xo = pnp_admm(Y[:, s], dict(param, F=Fo, net=box_denoiser, X0=X0[..., s]), solver="exact")
        assert rel_l2(x[..., s], xo) <= TOL_XUPDATE


def test_admm_iter_counts_zero_and_one(q, ops):
    P, Po = ops["spiral"]
    Fo, Xgt, Y, X0 = make_problem(Po, 23)
    for it in (0, 1):
        x = q.PnP_ADMM(Y, {"iter": it, "gamma": 0.05, "F": q.fft_operator(P), "net": box_denoiser, "X0": X0})
        assert rel_l2(x, X0) <= 1e-6  # x_1 = X0 because A X0 = y (PnP_ADMM.m:102 with a zero residual)


@pytest.mark.parametrize("dtype,precision", [("single_level", "fp32"), ("multi_level", "fp32"), ("single_level", "tc"), ("multi_level", "tc")])
def test_admm_loop_with_builtin_unetres(q, ops, dtype, precision):
    """End to end: the on-device loop (K1 + K3) against the oracle loop with the CPU UNetRes."""
    from oracle import unetres
    from oracle.admm import pnp_admm
    P, Po = ops["spiral"]
    Fo, Xgt, Y, X0 = make_problem(Po, 24)
    in_nc = 10 if dtype == "single_level" else 11
    sd = unetres.make_state_dict(in_nc, seed=0)
    net = q.UNetRes(sd, in_nc=in_nc)
    net.set_precision(precision)
    param = {"iter": 3, "gamma": 0.05, "X0": X0, "denoiser_type": dtype}
    if dtype == "multi_level":
        param["noise_map"] = q.build_noise_map(0.01, 224, 224)
    xo = pnp_admm(Y, dict(param, F=Fo, net=lambda v: unetres.denoise_matlab_layout(sd, v)), solver="exact")
    x = q.PnP_ADMM(Y, dict(param, F=q.fft_operator(P), net=net))
    assert rel_l2(x, xo) <= 1e-4  # bounded by the denoiser tolerance


def test_admm_graph_replay_equals_direct_launches(q, ops, monkeypatch):
    """Steady-state iterations are replayed from a captured CUDA graph; the result must be bit-identical to direct launches."""
    from oracle import unetres
    P, Po = ops["spiral"]
    Fo, Xgt, Y, X0 = make_problem(Po, 25, S=2)
    net = q.UNetRes(unetres.make_state_dict(10, seed=0), in_nc=10)
    param = {"iter": 8, "gamma": 0.05, "X0": X0, "denoiser_type": "single_level", "F": q.fft_operator(P), "net": net}
    ctx = q.Context.default()
    l0 = ctx.launch_count
    x_graph = q.PnP_ADMM(Y, param)
    n_graph = ctx.launch_count - l0
    monkeypatch.setenv("QMRI_NO_GRAPH", "1")
    l0 = ctx.launch_count
    x_direct = q.PnP_ADMM(Y, param)
    n_direct = ctx.launch_count - l0
    assert np.array_equal(x_graph, x_direct)
    assert n_graph == n_direct  # kernels inside the replayed graph are counted



# ---------------------------------------------------------------------------------------------
# K1 streaming kernel (one CTA per slice-channel; chosen automatically for large slice batches)
@pytest.fixture()
def stream_op(q, monkeypatch):
    from oracle import sampling
    monkeypatch.setenv("QMRI_K1_KERNEL", "stream")
    V = np.eye(10)
    return q.setup_subsampling_spiralgrided(224, 224, 771, V), sampling.setup_subsampling_spiralgrided(224, 224, 771, V)


def test_stream_kernel_operator_and_xupdate(q, stream_op):
    from oracle.sampling import FOperator
    from oracle.xupdate import xupdate_exact
    P, Po = stream_op
    F, Fo = q.fft_operator(P), FOperator(Po)
    x = smooth_tsmi(31, S=3, cplx=True)
    y = F.forward(x)
    for s in range(3):
        assert rel_l2(y[:, s], Fo.forward(x[..., s])) <= TOL_XUPDATE
    xr = smooth_tsmi(32)                       # real input (in_im == null path)
    assert rel_l2(F.forward(xr), Fo.forward(xr)) <= TOL_XUPDATE
    rng = np.random.default_rng(3)
    b = rng.standard_normal((P.nmeas, 3)) + 1j * rng.standard_normal((P.nmeas, 3))
    xa = F.adjoint(b)
    for s in range(3):
        assert rel_l2(xa[..., s], Fo.adjoint(b[:, s])) <= TOL_XUPDATE
    assert rel_l2(F.forward(xa), b) <= TOL_XUPDATE      # A A^H = I
    yv = Fo.forward(smooth_tsmi(33))
    v, u = smooth_tsmi(34), 0.1 * smooth_tsmi(35, cplx=True)
    xs, w, mm = F.xupdate(yv, v, u, 0.05, want_w=True)
    xo = xupdate_exact(Fo, yv, v - u, 0.05)
    assert rel_l2(xs, xo) <= TOL_XUPDATE
    wo = xo + u
    assert rel_l2(w, wo) <= TOL_XUPDATE
    assert abs(mm[0] - wo.real.min()) <= 1e-5 * abs(wo.real).max()
    assert abs(mm[1] - wo.real.max()) <= 1e-5 * abs(wo.real).max()


def test_stream_kernel_admm_loop_and_agreement_with_cluster_kernel(q, stream_op, monkeypatch):
    from oracle.admm import pnp_admm
    P, Po = stream_op
    monkeypatch.setenv("QMRI_K1_KERNEL", "cluster")
    Pc = q.setup_subsampling_spiralgrided(224, 224, 771, np.eye(10))
    Fo, Xgt, Y, X0 = make_problem(Po, 36, S=2)
    Y[:, 1] *= 2.0
    X0 = Fo.adjoint(Y)
    param = {"iter": 5, "gamma": 0.05, "denoiser_type": "single_level"}
    x = q.PnP_ADMM(Y, dict(param, F=q.fft_operator(P), net=box_denoiser, X0=X0))
    for s in range(2):
        xo = pnp_admm(Y[:, s], dict(param, F=Fo, net=box_denoiser, X0=X0[..., s]), solver="exact")
        assert rel_l2(x[..., s], xo) <= TOL_XUPDATE
    xc = q.PnP_ADMM(Y, dict(param, F=q.fft_operator(Pc), net=box_denoiser, X0=X0))
    assert rel_l2(x, xc) <= 2e-6


def test_stream_kernel_with_builtin_unetres_and_graph(q, stream_op):
    from oracle import unetres
    from oracle.admm import pnp_admm
    P, Po = stream_op
    Fo, Xgt, Y, X0 = make_problem(Po, 37)
    sd = unetres.make_state_dict(10, seed=0)
    net = q.UNetRes(sd, in_nc=10)
    param = {"iter": 4, "gamma": 0.05, "X0": X0, "denoiser_type": "single_level"}
    xo = pnp_admm(Y, dict(param, F=Fo, net=lambda v: unetres.denoise_matlab_layout(sd, v)), solver="exact")
    x = q.PnP_ADMM(Y, dict(param, F=q.fft_operator(P), net=net))
    assert rel_l2(x, xo) <= 1e-4


def test_stream_kernel_line_sampled_mask(q, monkeypatch):
    """EPI rows hold 224 samples each: they are cut into chunks whose partial sums meet in shared-memory overflow slots."""
    from oracle import sampling
    from oracle.sampling import FOperator
    from oracle.xupdate import xupdate_exact
    monkeypatch.setenv("QMRI_K1_KERNEL", "stream")
    V = np.eye(10)
    P, Po = q.setup_subsampling_epi(224, 224, 1 / 65, V), sampling.setup_subsampling_epi(224, 224, 1 / 65, V)
    F, Fo = q.fft_operator(P), FOperator(Po)
    x = smooth_tsmi(41, cplx=True)
    assert rel_l2(F.forward(x), Fo.forward(x)) <= TOL_XUPDATE
    rng = np.random.default_rng(4)
    b = rng.standard_normal(P.nmeas) + 1j * rng.standard_normal(P.nmeas)
    assert rel_l2(F.adjoint(b), Fo.adjoint(b)) <= TOL_XUPDATE
    y = Fo.forward(smooth_tsmi(42))
    v, u = smooth_tsmi(43), 0.1 * smooth_tsmi(44, cplx=True)
    assert rel_l2(F.xupdate(y, v, u, 0.05), xupdate_exact(Fo, y, v - u, 0.05)) <= TOL_XUPDATE


def test_cluster_kernel_batched(q, monkeypatch):
    """Batches of two slices or more default to the streaming kernels; the cluster kernel must stay correct for them too."""
    from oracle import sampling
    from oracle.sampling import FOperator
    from oracle.xupdate import xupdate_exact
    monkeypatch.setenv("QMRI_K1_KERNEL", "cluster")
    V = np.eye(10)
    P, Po = q.setup_subsampling_spiralgrided(224, 224, 771, V), sampling.setup_subsampling_spiralgrided(224, 224, 771, V)
    F, Fo = q.fft_operator(P), FOperator(Po)
    x = smooth_tsmi(51, S=3, cplx=True)
    y = F.forward(x)
    for s in range(3):
        assert rel_l2(y[:, s], Fo.forward(x[..., s])) <= TOL_XUPDATE
    yv = Fo.forward(smooth_tsmi(52))
    v, u = smooth_tsmi(53), 0.1 * smooth_tsmi(54, cplx=True)
    assert rel_l2(F.xupdate(yv, v, u, 0.05), xupdate_exact(Fo, yv, v - u, 0.05)) <= TOL_XUPDATE


def test_large_batch_takes_streaming_kernel_and_matches_oracle(q, ops):
    """Two slices or more take the streaming kernels on their own: 32 slices (one short last wave), spot-check four of them."""
    from oracle.sampling import FOperator
    from oracle.xupdate import xupdate_exact
    P, Po = ops["spiral"]
    F, Fo = q.fft_operator(P), FOperator(Po)
    S = 32
    rng = np.random.default_rng(5)
    base = smooth_tsmi(38, cplx=True)
    scale = 1.0 + rng.random(S)
    x = base[..., None] * scale
    y = F.forward(x)
    assert y.shape == (P.nmeas, S)
    yo = Fo.forward(base)
    for s in (0, 7, 19, 31):
        assert rel_l2(y[:, s], yo * scale[s]) <= TOL_XUPDATE


# ---------------------------------------------------------------------------------------------
# general V: frame i samples sum_c V(i,c) X^_c on its own mask; exact block solve per k-space location
def general_ops(q, kind, V):
    from oracle import sampling
    if kind == "spiral":
        return q.setup_subsampling_spiralgrided(224, 224, 771, V), sampling.setup_subsampling_spiralgrided(224, 224, 771, V)
    return q.setup_subsampling_epi(224, 224, 1 / 65, V), sampling.setup_subsampling_epi(224, 224, 1 / 65, V)


def tsmi_c(seed, C, S=None, cplx=False):
    x = smooth_tsmi(seed, S=S, cplx=cplx)
    return x[:, :, :C] if S is None else x[:, :, :C, :]


@pytest.mark.parametrize("kind,L,C", [("spiral", 12, 10), ("spiral", 6, 4), ("epi", 5, 4)])
def test_general_V_operator_and_xupdate(q, kind, L, C):
    from oracle.sampling import FOperator
    from oracle.xupdate import xupdate_exact
    rng = np.random.default_rng(60 + L)
    V = rng.standard_normal((L, C)) / np.sqrt(C)
    P, Po = general_ops(q, kind, V)
    F, Fo = q.fft_operator(P), FOperator(Po)
    assert P.nmeas == Po.nmeas
    x = tsmi_c(61, C, S=2, cplx=True)
    y = F.forward(x)
    for s in range(2):
        assert rel_l2(y[:, s], Fo.forward(x[..., s])) <= TOL_XUPDATE
    b = rng.standard_normal((P.nmeas, 2)) + 1j * rng.standard_normal((P.nmeas, 2))
    xa = F.adjoint(b)
    for s in range(2):
        assert rel_l2(xa[..., s], Fo.adjoint(b[:, s])) <= TOL_XUPDATE
    lhs, rhs = np.vdot(b, y), np.vdot(xa, x)             # <A x, b> == <x, A^H b>
    assert abs(lhs - rhs) <= 1e-5 * abs(lhs)
    yv = Fo.forward(tsmi_c(62, C))
    v, u = tsmi_c(63, C), 0.1 * tsmi_c(64, C, cplx=True)
    for rho in (0.05, 0.7):
        xs, w, mm = F.xupdate(yv, v, u, rho, want_w=True)
        xo = xupdate_exact(Fo, yv, v - u, rho)
        assert rel_l2(xs, xo) <= TOL_XUPDATE
        assert rel_l2(w, xo + u) <= TOL_XUPDATE


@pytest.mark.parametrize("L,C", [(40, 4), (200, 10)])
def test_general_V_many_frames(q, L, C):
    """real(dict.V) with T rows: the union of 200 spiral masks holds 11 051 locations and is processed in several parts
    (one forward launch per part, accumulating adjoint passes)."""
    from oracle.sampling import FOperator
    from oracle.xupdate import xupdate_exact
    rng = np.random.default_rng(80 + L)
    V = np.linalg.qr(rng.standard_normal((L, C)))[0]               # orthonormal columns, like an SVD subspace
    P, Po = general_ops(q, "spiral", V)
    F, Fo = q.fft_operator(P), FOperator(Po)
    assert P.nmeas == Po.nmeas
    x = tsmi_c(81, C, cplx=True)
    y = F.forward(x)
    yo = Fo.forward(x)
    assert rel_l2(y, yo) <= TOL_XUPDATE
    b = rng.standard_normal(P.nmeas) + 1j * rng.standard_normal(P.nmeas)
    assert rel_l2(F.adjoint(b), Fo.adjoint(b)) <= TOL_XUPDATE
    v, u = tsmi_c(82, C), 0.1 * tsmi_c(83, C, cplx=True)
    xs, w, mm = F.xupdate(yo, v, u, 0.05, want_w=True)
    xo = xupdate_exact(Fo, yo, v - u, 0.05)
    assert rel_l2(xs, xo) <= TOL_XUPDATE
    wo = xo + u
    assert abs(mm[0] - wo.real.min()) <= 1e-5 * abs(wo.real).max() and abs(mm[1] - wo.real.max()) <= 1e-5 * abs(wo.real).max()


def test_general_V_admm_loop(q):
    from oracle.admm import pnp_admm
    from oracle.sampling import FOperator
    from oracle.synth import awgn_measured
    rng = np.random.default_rng(70)
    V = np.linalg.qr(rng.standard_normal((24, 10)))[0] * 1.3      # 24 frames (union in two parts), 10 channels, A A^H != I
    P, Po = general_ops(q, "spiral", V)
    Fo = FOperator(Po)
    Xgt = smooth_tsmi(71, S=2)
    Y = np.stack([awgn_measured(Fo.forward(Xgt[..., s]), 30, 71 + s) for s in range(2)], axis=1)
    X0 = Fo.adjoint(Y)
    param = {"iter": 5, "gamma": 0.05, "denoiser_type": "single_level"}
    x = q.PnP_ADMM(Y, dict(param, F=q.fft_operator(P), net=box_denoiser, X0=X0))
    for s in range(2):
        xo = pnp_admm(Y[:, s], dict(param, F=Fo, net=box_denoiser, X0=X0[..., s]), solver="exact")
        assert rel_l2(x[..., s], xo) <= TOL_XUPDATE


def test_admm_honours_any_X0(q, ops):
    """param.X0 is the caller's: iteration 1 solves from it (PnP_ADMM.m:76-78,102), whether or not it equals A^H y."""
    from oracle.admm import pnp_admm
    P, Po = ops["spiral"]
    Fo, Xgt, Y, X0 = make_problem(Po, 72)
    X0b = 0.5 * X0 + 0.1 * smooth_tsmi(73, cplx=True)
    for it in (1, 3):
        param = {"iter": it, "gamma": 0.05, "denoiser_type": "single_level"}
        xo = pnp_admm(Y, dict(param, F=Fo, net=box_denoiser, X0=X0b), solver="exact")
        x = q.PnP_ADMM(Y, dict(param, F=q.fft_operator(P), net=box_denoiser, X0=X0b))
        assert rel_l2(x, xo) <= TOL_XUPDATE


# ---------------------------------------------------------------------------------------------
# full size (BASELINE configs[3]: 120 slices): size-independent properties
def test_full_size_120_slices_recon_and_matching(q, ops):
    """120-slice PnP-ADMM batch with the built-in UNetRes + matching of every pixel.  Slices are independent, so (a) a slice of the
    batch equals the same slice reconstructed alone, (b) a permuted batch gives the permuted result (to the denoiser tolerance: kernel shapes follow the chunk), (c) two slices
    against the oracle loop, (d) matched indices of the whole batch equal those of the same slices matched alone."""
    from oracle import synth, unetres
    from oracle.admm import pnp_admm
    import bench
    P, Po = ops["spiral"]
    F = q.fft_operator(P)
    S = 120
    X = bench.synthetic_slices(S, seed=77)
    Y = F.forward(X)
    rng = np.random.default_rng(78)
    Y = Y + (rng.standard_normal(Y.shape) + 1j * rng.standard_normal(Y.shape)) * np.sqrt(np.mean(np.abs(Y) ** 2) / 10 ** 3.0 / 2)
    X0 = F.adjoint(Y)
    sd = unetres.make_state_dict(10, seed=0)
    net = q.UNetRes(sd, in_nc=10)
    param = {"iter": 6, "gamma": 0.05, "denoiser_type": "single_level", "F": F, "net": net}
    x = q.PnP_ADMM(Y, dict(param, X0=X0))
    assert x.shape == (224, 224, 10, S) and np.isfinite(x).all()
    pick = [0, 59, 119]
    for s in pick:                                                     # (a)
        xs = q.PnP_ADMM(Y[:, s], dict(param, X0=X0[..., s]))
        assert rel_l2(x[..., s], xs) <= TOL_DENOISER                   # tile / split-K / x-update kernel choices depend on the batch size: not bitwise
    perm = rng.permutation(S)                                          # (b)
    xp = q.PnP_ADMM(np.ascontiguousarray(Y[:, perm]), dict(param, X0=np.ascontiguousarray(X0[..., perm])))
    assert rel_l2(xp, x[..., perm]) <= TOL_DENOISER       # slices sit at other positions of the denoiser pass (tile shapes / split-K follow the position): not bitwise
    assert np.isfinite(xp).all()
    from oracle.sampling import FOperator                              # (c)
    Fo = FOperator(Po)
    for s in pick[:2]:
        xo = pnp_admm(Y[:, s], dict(param, F=Fo, net=lambda v: unetres.denoise_matlab_layout(sd, v), X0=X0[..., s]), solver="exact")
        assert rel_l2(x[..., s], xo) <= 1e-4
    d = synth.make_dictionary(K_target=20000, cut=3, seed=0)          # (d)
    out = q.mrf_dtm_cpu(d, {"X": np.transpose(x, (0, 1, 3, 2))}, {"f": {"qout": 1, "pdout": 1, "dmout": 1}})   # [Nx, Ny, Nz, T]
    assert out["dm"].shape == (224, 224, S) and out["qmap"].shape == (224, 224, S, 2)
    for s in pick:
        o1 = q.mrf_dtm_cpu(d, {"X": x[..., s]}, {"f": {"qout": 1, "pdout": 1, "dmout": 1}})
        assert np.array_equal(o1["dm"], out["dm"][:, :, s]) and np.array_equal(o1["qmap"], out["qmap"][:, :, s, :])

# ---------------------------------------------------------------------------------------------
def test_error_behaviour(q):
    V = np.eye(10)
    with pytest.raises(q.QmriError):
        q.setup_subsampling_spiralgrided(224, 224, 771, np.ones((12, 20)))       # general V: at most 16 channels
    with pytest.raises(q.QmriError):
        q.setup_subsampling_spiralgrided(128, 128, 771, V)                       # unsupported size
    with pytest.raises(q.QmriError):
        q.setup_subsampling_epi(224, 224, 0.0, V)
    P = q.setup_subsampling_epi(224, 224, 1 / 65, V)
    F = q.fft_operator(P)
    with pytest.raises(ValueError):
        F.forward(np.zeros((224, 224, 9)))
    with pytest.raises(KeyError):
        q.PnP_ADMM(np.zeros(P.nmeas, complex), {"iter": 1, "gamma": 0.05, "F": F, "X0": np.zeros((224, 224, 10))})
    with pytest.raises(q.QmriError):
        q.PnP_ADMM(np.zeros(P.nmeas, complex), {"iter": 1, "gamma": 0.0, "F": F, "X0": np.zeros((224, 224, 10)), "net": box_denoiser})

    def bad(v):
        raise RuntimeError("denoiser exploded")
    with pytest.raises(RuntimeError, match="exploded"):
        q.PnP_ADMM(np.zeros(P.nmeas, complex), {"iter": 2, "gamma": 0.05, "F": F, "X0": np.ones((224, 224, 10)), "net": bad})
