"""GPU parity tests of the rows either side of the loop: P.for / P.adj, measurement noise, foreground mask and
metrics, trained-weight import, cut0-cut4 dictionaries, shard-resident dictionaries - CUDA path (through the C ABI) vs the oracle."""
import numpy as np
import pytest

from conftest import rel_l2

pytestmark = pytest.mark.gpu

NEAR_TIE = 1e-6


@pytest.fixture(scope="module")
def q():
    import qmri_b200
    qmri_b200.Context.default()
    return qmri_b200


# ---- P.for / P.adj (setup_subsampling_spiralgrided.m:41-42, setup_subsampling_epi.m:34-35) -----------------------------------
@pytest.mark.parametrize("kind", ["spiral", "epi"])
@pytest.mark.parametrize("general", [False, True])
def test_p_for_adj_match_the_sparse_matrix(q, kind, general):
    from oracle import sampling
    rng = np.random.default_rng(3)
    V = rng.standard_normal((12, 10)) / 3 if general else np.eye(10)
    if kind == "spiral":
        P, Po = q.setup_subsampling_spiralgrided(224, 224, 771, V), sampling.setup_subsampling_spiralgrided(224, 224, 771, V)
    else:
        P, Po = q.setup_subsampling_epi(224, 224, 1 / 65, V), sampling.setup_subsampling_epi(224, 224, 1 / 65, V)
    n = 224 * 224 * 10
    x = rng.standard_normal(n) + 1j * rng.standard_normal(n)
    y = rng.standard_normal(Po.nmeas) + 1j * rng.standard_normal(Po.nmeas)
    yf, xa = P.for_(x), P["adj"](y)
    if general:
        assert rel_l2(yf, Po.for_(x)) < 1e-14 and rel_l2(xa, Po.adj(y)) < 1e-14
    else:
        assert np.array_equal(yf, Po.for_(x)) and np.array_equal(xa, Po.adj(y))     # a pure gather / scatter in double
    # the reference composes F from them (main_recon_tsmis_FFT.m:228-229): same result as the fused operator
    X = rng.standard_normal((224, 224, 10))
    F = q.fft_operator(P)
    y_ref = P.for_(np.fft.fft2(X, axes=(0, 1)).reshape(-1, order="F")) / 224.0
    assert rel_l2(F.forward(X), y_ref) < 1e-5
    x_ref = np.fft.ifft2(P.adj(y).reshape((224, 224, 10), order="F"), axes=(0, 1)) * 224.0
    assert rel_l2(F.adjoint(y), x_ref) < 1e-5
    # single precision in, single precision out
    assert P.for_(x.astype(np.complex64)).dtype == np.complex64
    with pytest.raises(ValueError):
        P.for_(x[:-1])


# ---- awgn(Y, snr, 'measured') (main_recon_tsmis_FFT.m:243) ---------------------------------------------------------------------
def test_awgn_matches_its_restatement_and_the_statistics(q):
    from oracle.metrics import awgn_philox
    rng = np.random.default_rng(0)
    nmeas, S = 6184, 5
    Y = (rng.standard_normal((nmeas, S)) + 1j * rng.standard_normal((nmeas, S))) * np.array([1.0, 2.0, 0.5, 3.0, 1.5])
    Yn = q.awgn(Y, 30.0, "measured", seed=1234)
    assert Yn.shape == Y.shape and Yn.dtype == np.complex128
    assert rel_l2(Yn - Y, awgn_philox(Y, 30.0, 1234) - Y) < 1e-10          # same generator, same Box-Muller, double throughout
    noise = Yn - Y
    for s in range(S):
        p_sig = np.mean(np.abs(Y[:, s]) ** 2)
        p_n = np.mean(np.abs(noise[:, s]) ** 2)
        assert abs(10 * np.log10(p_sig / p_n) - 30.0) < 0.3                 # 'measured': per-call signal power; 6184 samples: +-0.11 dB (2 sigma)
        assert abs(np.mean(noise[:, s])) < 4 * np.sqrt(p_n / nmeas)
        # circular: real and imaginary parts carry equal power and are uncorrelated
        assert abs(np.var(noise[:, s].real) / np.var(noise[:, s].imag) - 1) < 0.1
        assert abs(np.corrcoef(noise[:, s].real, noise[:, s].imag)[0, 1]) < 0.06
    # Gaussian: kurtosis of the real part ~ 3
    z = noise[:, 0].real / noise[:, 0].real.std()
    assert abs(np.mean(z ** 4) - 3.0) < 0.25
    # reproducible, seed-dependent, independent of the batch the slice sits in, input untouched
    assert np.array_equal(Yn, q.awgn(Y, 30.0, "measured", seed=1234))
    assert not np.array_equal(Yn, q.awgn(Y, 30.0, "measured", seed=1235))
    assert np.array_equal(q.awgn(Y[:, :2], 30.0, "measured", seed=1234), Yn[:, :2])
    y32 = q.awgn(Y[:, 0].astype(np.complex64), 20.0, "measured", seed=7)
    assert y32.dtype == np.complex64 and y32.shape == (nmeas,)
    assert abs(10 * np.log10(np.mean(np.abs(Y[:, 0]) ** 2) / np.mean(np.abs(y32 - Y[:, 0]) ** 2)) - 20.0) < 0.3
    with pytest.raises(ValueError):
        q.awgn(Y, 30.0, "other")


# ---- getmask_fromPD + metrics (main_recon_tsmis_FFT.m:190, 328-384) ---------------------------------------------------------------
def test_foreground_mask_matches_oracle(q):
    import benchdata
    from oracle import metrics
    qm = benchdata.volunteer_slices(0, 15)
    for s in (0, 7, 14):
        pd = qm[s, 2]
        assert np.array_equal(q.getmask_fromPD(pd, 0.15), metrics.getmask_fromPD(pd, 0.15))
    # crafted cases: nested rings, a diagonal leak, a spiral corridor (many flood iterations), complex input
    rng = np.random.default_rng(1)
    img = np.zeros((64, 48))
    img[4:60, 4:44] = 1.0
    img[8:56, 8:40] = 0.0
    img[12:52, 12:36] = 0.7
    img[20:40, 18:30] = 0.05                # below threshold inside the inner block: a hole -> filled
    img[4, 4] = 0.0
    img[5, 5] = 0.0
    img[6, 6] = 0.0
    img[7, 7] = 0.0                          # diagonal corridor through the outer ring: 8-connected background leaks in
    assert np.array_equal(q.getmask_fromPD(img, 0.15), metrics.getmask_fromPD(img, 0.15))
    spiral = np.ones((41, 41))
    r, c, dr, dc, run = 20, 20, 0, 1, 1
    spiral[r, c] = 0
    while True:                               # carve a one-pixel spiral corridor out to the border
        done = False
        for _ in range(2):
            for _ in range(run):
                r, c = r + dr, c + dc
                if not (0 <= r < 41 and 0 <= c < 41):
                    done = True
                    break
                spiral[r, c] = 0
            if done:
                break
            dr, dc = dc, -dr
        if done:
            break
        run += 2
    assert np.array_equal(q.getmask_fromPD(spiral, 0.5), metrics.getmask_fromPD(spiral, 0.5))
    cplx = (rng.random((32, 32)) > 0.4) * np.exp(1j * rng.random((32, 32)))
    assert np.array_equal(q.getmask_fromPD(cplx, 0.15), metrics.getmask_fromPD(cplx, 0.15))


def test_recon_metrics_match_oracle(q):
    import benchdata
    from oracle import metrics
    rng = np.random.default_rng(2)
    qm0 = np.transpose(benchdata.volunteer_slices(3, 4)[0], (1, 2, 0))                 # N x M x 3 ground truth
    mask = metrics.getmask_fromPD(qm0[:, :, 2], 0.15)
    # "reconstructed" maps: perturbed T1 / T2 and a complex PD with an arbitrary scale (the script normalises |PD| by its max)
    est = np.empty((224, 224, 3), np.complex64)
    est[:, :, 0] = qm0[:, :, 0] * (1 + 0.05 * rng.standard_normal((224, 224)))
    est[:, :, 1] = qm0[:, :, 1] * (1 + 0.08 * rng.standard_normal((224, 224)))
    est[:, :, 2] = 37.0 * qm0[:, :, 2] * (1 + 0.03 * rng.standard_normal((224, 224))) * np.exp(1j * 0.3)
    X0 = rng.standard_normal((224, 224, 10)) * mask[:, :, None] * 0.2
    X = X0 + 0.01 * (rng.standard_normal(X0.shape) + 1j * rng.standard_normal(X0.shape))
    got = q.recon_metrics(est, qm0, mask, X, X0)
    ref = metrics.recon_metrics(est, qm0, mask, X, X0)
    assert set(got) == set(ref) and len(got) == 11
    for k in ref:
        assert got[k] == pytest.approx(ref[k], rel=1e-9, abs=1e-12), k
    # no mask, no TSMIs, real double inputs
    got2 = q.recon_metrics(np.real(est).astype(np.float64), qm0)
    ref2 = metrics.recon_metrics(np.real(est).astype(np.float64), qm0, None)
    for k in ref2:
        assert got2[k] == pytest.approx(ref2[k], rel=1e-9, abs=1e-12), k
    assert np.isnan(got2["tsmi_mean_psnr"])
    # identical images: MAE 0, PSNR Inf, SSIM 1
    same = q.recon_metrics(qm0, qm0, mask, X0, X0)
    assert same["t1_mae"] == 0 and np.isinf(same["t2_psnr"]) and same["pd_ssim"] == pytest.approx(1.0, abs=1e-12)
    assert np.isinf(same["tsmi_mean_psnr"]) and same["tsmi_mean_ssim"] == pytest.approx(1.0, abs=1e-12)


# ---- trained weights: .pt checkpoint and ONNX initialisers -> the same forward as the CPU network ------------------------------
@pytest.mark.parametrize("route", ["pt", "onnx"])
def test_trained_weight_import_runs_on_the_gpu(q, route, tmp_path):
    import torch
    from oracle import unetres
    sd = unetres.make_state_dict(11, seed=9)
    if route == "pt":
        f = tmp_path / "ckpt.pt"
        torch.save({"model_state_dict": sd, "epoch": 3, "loss": 0.5}, f)       # main_train.py's format, read at main_test.py:260-262
        net = q.UNetRes.from_checkpoint(str(f))
    else:
        from test_host_logic import _onnx_model
        keys = q.state_dict_keys(11)
        tensors = [(k, sd[k].numpy(), True, True) for k, _ in keys]
        nodes = [("ConvTranspose" if k.startswith("m_up") and k.endswith(".0.weight") else "Conv", ["a", k]) for k, _ in keys]
        f = tmp_path / "net.onnx"
        f.write_bytes(_onnx_model(tensors, nodes))
        net = q.UNetRes.from_onnx(str(f))
    assert net.in_nc == 11
    rng = np.random.default_rng(4)
    x = rng.random((1, 11, 64, 48)).astype(np.float32)
    with torch.no_grad():
        ref = unetres.unetres_forward(sd, torch.from_numpy(x)).numpy()
    for mode in ("fp32", "tc"):
        net.set_precision(mode)
        assert rel_l2(net.forward(x), ref) <= 1e-4


# ---- cut0 .. cut4 (main_recon_tsmis_FFT.m:42: T = 1000, 500, 300, 200, 100) ----------------------------------------------------------
@pytest.mark.parametrize("cut", [0, 1, 2, 3, 4])
def test_matching_and_synthesis_over_all_cuts(q, cut):
    import benchdata
    from oracle import synth
    from oracle.matching import mrf_dtm_cpu as oracle_match
    d = benchdata.make_dictionary(K_target=6000, cut=cut, seed=cut)
    assert d["V"].shape == (benchdata.CUT_T[cut], 10)
    qm = benchdata.volunteer_slices(20 + cut, 21 + cut)[0][:, 60:124, 70:134]        # a 64 x 64 window of a slice
    X, I = q.synthesize_tsmis(d, qm, return_index=True)
    Xo, Io = synth.synthesize_tsmis(d, qm)
    lut = d["lut"].astype(np.float64)
    qv = np.transpose(qm, (1, 2, 0)).reshape((-1, 3), order="F")
    dg = np.sum((lut[I] - qv[:, :2]) ** 2, axis=1)
    do = np.sum((lut[Io] - qv[:, :2]) ** 2, axis=1)
    same = I == Io
    assert np.all(same | (np.abs(dg - do) <= 1e-6 * np.maximum(do, 1e-12)))          # different atom only at (fp32) distance ties
    assert rel_l2(X[same.reshape((64, 64), order="F")], Xo[same.reshape((64, 64), order="F")]) < 1e-6
    rng = np.random.default_rng(cut)
    Xn = X + 0.02 * np.abs(X).max() * (rng.standard_normal(X.shape) + 1j * rng.standard_normal(X.shape))
    out = q.mrf_dtm_cpu(d, {"X": Xn}, {"f": {"qout": 1, "pdout": 1, "dmout": 1, "mtout": 1}})
    ref = oracle_match(d, {"X": Xn}, None, return_gap=True)
    decided = ref["gap"] >= NEAR_TIE
    assert decided.mean() > 0.5                                                      # smooth dictionaries have many near-ties
    assert np.array_equal(out["dm"].astype(np.int64)[decided], ref["dm"][decided])
    assert np.array_equal(out["qmap"][decided], ref["qmap"][decided])
    assert rel_l2(out["pd"][decided], ref["pd"][decided]) < 1e-5 and rel_l2(out["mt"][decided], ref["mt"][decided]) < 1e-5


# ---- shard-resident dictionary (BASELINE config 5): two "ranks" on one GPU -------------------------------------------------------
def test_shard_resident_dictionary_equals_whole(q):
    import ctypes as C
    import torch
    import benchdata
    d = benchdata.make_dictionary(K_target=30000, cut=3, seed=0)
    K = d["D"].shape[0]
    whole = q.Dictionary(d)
    npix = 5000
    g = torch.Generator(device="cuda").manual_seed(1)
    xr = torch.randn(10 * npix, device="cuda", generator=g)
    xi = torch.randn(10 * npix, device="cuda", generator=g)
    # plant exact cross-shard ties: pixel p's signature = an atom of shard 0; a copy of that atom also sits in shard 1
    d2 = dict(d)
    D = d["D"].copy()
    D[K - 5] = D[3]
    d2["D"] = D
    whole = q.Dictionary(d2)
    xr.view(10, npix)[:, 0] = torch.from_numpy(D[3]).cuda()
    xi.view(10, npix)[:, 0] = 0
    ref = q.mrf_dtm_sharded(whole, xr, xi, npix, want_mt=True)
    outs, keys = [], []
    cuts = [0, K // 3, K]
    shards = []
    for r in range(2):
        a0, a1 = cuts[r], cuts[r + 1]
        part = dict(d2, D=D[a0:a1])
        shards.append(q.Dictionary(part, shard=(a0, a1), shard_only=True))
        k = torch.zeros(npix, dtype=torch.int64, device="cuda")
        torch.cuda.synchronize()
        q._capi.check(whole.ctx.lib.qmri_match_keys_dev(shards[r].handle, C.c_void_p(xr.data_ptr()), C.c_void_p(xi.data_ptr()), npix, C.c_void_p(k.data_ptr())))
        whole.ctx.synchronize()
        keys.append(k)
    merged = torch.maximum(keys[0], keys[1])                                  # what ncclAllReduce(max) computes
    tot = None
    for r in range(2):
        buf = torch.empty(6 * npix, dtype=torch.float32, device="cuda")
        dm = torch.empty(npix, dtype=torch.int32, device="cuda")
        torch.cuda.synchronize()
        q._capi.check(whole.ctx.lib.qmri_match_finish_dev(shards[r].handle, C.c_void_p(xr.data_ptr()), C.c_void_p(xi.data_ptr()), npix,
                                                        C.c_void_p(merged.data_ptr()), C.c_void_p(buf.data_ptr()), C.c_void_p(buf.data_ptr() + 8 * npix),
                                                        C.c_void_p(buf.data_ptr() + 16 * npix), C.c_void_p(dm.data_ptr())))
        whole.ctx.synchronize()
        part = torch.cat([buf[:5 * npix], dm.to(torch.float32)])
        tot = part if tot is None else tot + part                             # what ncclAllReduce(sum) computes
    qmap, pd, mt, dm = ref
    assert torch.equal(tot[:2 * npix].view(2, npix), qmap)
    assert torch.equal(tot[2 * npix:4 * npix].view(npix, 2), pd)
    assert torch.equal(tot[4 * npix:5 * npix], mt)
    assert torch.equal(tot[5 * npix:].to(torch.int32), dm)
    assert int(dm[0]) == 4                                                    # tie across shards: the lower atom index (3, 1-based 4) wins
    with pytest.raises(q.QmriError):
        q.mrf_dtm_cpu(shards[1], {"X": np.zeros((4, 4, 10))})                 # host entry point needs the whole dictionary
    for s in shards:
        s.close()
    whole.close()


# ---- LRTV baseline (FISTA_deep.m + prox_tv; main_recon_tsmis_FFT.m:272-282) ------------------------------------------------------
@pytest.mark.parametrize("kind", ["spiral", "epi"])
def test_lrtv_matches_oracle(q, kind):
    import bench
    from oracle import lrtv, sampling
    V = np.eye(10)
    if kind == "spiral":
        P, Po = q.setup_subsampling_spiralgrided(224, 224, 771, V), sampling.setup_subsampling_spiralgrided(224, 224, 771, V)
    else:
        P, Po = q.setup_subsampling_epi(224, 224, 1 / 65, V), sampling.setup_subsampling_epi(224, 224, 1 / 65, V)
    F, Fo = q.fft_operator(P), sampling.FOperator(Po)
    X0 = bench.synthetic_slices(1, 11)[..., 0]
    Y = q.awgn(Fo.forward(X0), 30.0, "measured", seed=3)
    step0 = X0.size / Y.size                                             # param.step = numel(X0)/numel(Y)
    param = {"K": 4e-5, "iter": 5, "step": step0, "tol": 1e-4, "backtrack": 1}
    data = {"N": 224, "M": 224, "L": 10, "y": Y, "F": F, "D": []}
    x, info = q.FISTA_deep(data, param, return_info=True)
    trace = []
    xo, its = lrtv.fista_lrtv(Y, Fo, X0.shape, K=4e-5, max_iter=5, step=step0, tol=1e-4, backtrack=True, trace=trace)
    assert info["iter"] == its and info["step"] == trace[-1]["step"]      # same backtracking decisions (step halved 7 times from 81)
    assert x.shape == (224, 224, 10) and x.dtype == np.complex128
    assert rel_l2(x, xo) < 1e-4                                           # operator in fp32, TV prox in double on both sides
    # K = 0: plain FISTA on the data term, no prox
    x0, _ = q.FISTA_deep(data, dict(param, K=0.0, iter=3), return_info=True)
    xo0, _ = lrtv.fista_lrtv(Y, Fo, X0.shape, K=0.0, max_iter=3, step=step0, tol=1e-4, backtrack=True)
    assert rel_l2(x0, xo0) < 1e-5
    with pytest.raises(KeyError):
        q.FISTA_deep(data, {"K": 1e-5, "iter": 2})
