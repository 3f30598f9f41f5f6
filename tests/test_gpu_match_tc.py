"""K2 tensor-pipe variant (tcgen05 kind::tf32, split operands, top-4 + FP32 rescore) against the oracle and against the FP32-FMA
kernel: identical atom indices outside oracle near-ties (top-2 relative gap < 1e-6), bit-identical outputs for the same index,
MATLAB's first-index rule on exact ties, NaN / zero pixels, ragged pixel counts, atom ranges that are not tile multiples."""
import ctypes as C

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


def _pixels(d, npix, cplx, seed, noise=0.02):
    rng = np.random.default_rng(seed)
    D = d["D"].astype(np.float64)
    idx = rng.integers(0, D.shape[0], npix)
    X = D[idx] * d["normD"][idx, None] * rng.uniform(0.3, 1.0, (npix, 1))
    if cplx:
        X = X * np.exp(1j * rng.uniform(0, 2 * np.pi, (npix, 1)))
        X = X + noise * np.abs(X).max() * (rng.standard_normal(X.shape) + 1j * rng.standard_normal(X.shape))
    else:
        X = X + noise * np.abs(X).max() * rng.standard_normal(X.shape)
    return X.reshape((npix, 1, D.shape[1]))


@pytest.mark.parametrize("cplx", [True, False])
@pytest.mark.parametrize("K,Cc,npix", [(6000, 10, 5000), (30000, 10, 4097), (3000, 4, 1000), (2500, 8, 777)])
def test_tensor_pipe_matches_oracle_and_fma(q, monkeypatch, cplx, K, Cc, npix):
    import benchdata
    from oracle.matching import mrf_dtm_cpu as oracle_match
    d = benchdata.make_dictionary(K_target=K, cut=3, C=Cc, seed=1)
    X = _pixels(d, npix, cplx, seed=K + npix)
    par = {"f": {"qout": 1, "pdout": 1, "dmout": 1, "mtout": 1}}
    ref = oracle_match(d, {"X": X}, None, return_gap=True)
    decided = ref["gap"] >= NEAR_TIE
    outs = {}
    ctx = q.Context.default()
    for pipe in ("tensor", "fma"):
        monkeypatch.setenv("QMRI_K2_PIPE", pipe)
        l0 = ctx.launch_count
        outs[pipe] = q.mrf_dtm_cpu(d, {"X": X}, par)
        assert ctx.launch_count - l0 == (4 if pipe == "tensor" else 3)        # unpack, [prep,] keys, finish
        dm = outs[pipe]["dm"].astype(np.int64)
        assert np.array_equal(dm[decided], ref["dm"][decided]), pipe
        assert np.array_equal(outs[pipe]["qmap"][decided], ref["qmap"][decided]), pipe
        assert rel_l2(outs[pipe]["pd"][decided], ref["pd"][decided]) < 1e-5 and rel_l2(outs[pipe]["mt"][decided], ref["mt"][decided]) < 1e-5
    # the two pipes: same index -> bit-identical outputs (the finish kernel is shared); different index only at near-ties
    same = outs["tensor"]["dm"] == outs["fma"]["dm"]
    assert same[decided].all()
    assert same.mean() > 0.97
    for k in ("mt", "pd", "qmap"):
        assert np.array_equal(outs["tensor"][k][same], outs["fma"][k][same])
    # where they differ the fp32 scores are within rounding of each other
    if (~same).any():
        assert np.max(np.abs(outs["tensor"]["mt"][~same] - outs["fma"]["mt"][~same]) / outs["fma"]["mt"][~same]) < 2e-6


def test_tensor_pipe_ties_nan_zero_and_ranges(q, monkeypatch):
    import torch
    import benchdata
    monkeypatch.setenv("QMRI_K2_PIPE", "tensor")
    d = benchdata.make_dictionary(K_target=5000, cut=3, seed=2)
    K = d["D"].shape[0]
    D = d["D"].copy()
    D[4000] = D[17]                      # exact duplicate atoms far apart (different tiles): the first index must win
    D[18] = D[17]                        # and adjacent
    d = dict(d, D=D)
    npix = 300
    X = _pixels(d, npix, True, seed=3)
    X[0, 0, :] = D[17] * (2.0 + 1.0j)    # exactly on the duplicated atom
    X[1, 0, :] = 0.0                     # zero pixel: every score 0 -> atom 1 (MATLAB's max of equal values)
    X[2, 0, :] = np.nan                  # NaN pixel: max(abs(NaN)) -> index 1
    X[3, 0, 0] = np.nan                  # partially NaN: still all scores NaN
    out = q.mrf_dtm_cpu(d, {"X": X}, {"f": {"qout": 1, "pdout": 1, "dmout": 1}})
    dm = out["dm"].astype(np.int64)[:, 0]
    assert dm[0] == 18 and dm[1] == 1 and dm[2] == 1 and dm[3] == 1
    monkeypatch.setenv("QMRI_K2_PIPE", "fma")
    out_f = q.mrf_dtm_cpu(d, {"X": X}, {"f": {"qout": 1, "pdout": 1, "dmout": 1}})
    assert np.array_equal(out_f["dm"][:4], out["dm"][:4])
    # atom ranges that are not multiples of the 128-atom tile: three handles, keys merged by max = the whole dictionary
    monkeypatch.setenv("QMRI_K2_PIPE", "tensor")
    whole = q.Dictionary(d)
    xs = np.asfortranarray(X.reshape((npix, 10)).astype(np.complex64))
    xs[2:4] = xs[5:7]                    # no NaN rows here
    xr = torch.from_numpy(np.ascontiguousarray(xs.real.T)).cuda().reshape(-1)
    xi = torch.from_numpy(np.ascontiguousarray(xs.imag.T)).cuda().reshape(-1)
    ref = q.mrf_dtm_sharded(whole, xr, xi, npix, want_mt=True)
    cuts = [0, 2049, 4003, K]
    merged = torch.zeros(npix, dtype=torch.int64, device="cuda")
    parts = []
    for r in range(3):
        h = q.Dictionary(d, shard=(cuts[r], cuts[r + 1]))
        parts.append(h)
        k = torch.zeros(npix, dtype=torch.int64, device="cuda")
        torch.cuda.synchronize()
        q._capi.check(whole.ctx.lib.qmri_match_keys_dev(h.handle, C.c_void_p(xr.data_ptr()), C.c_void_p(xi.data_ptr()), npix, C.c_void_p(k.data_ptr())))
        whole.ctx.synchronize()
        merged = torch.maximum(merged, k)
    qmap = torch.empty(2 * npix, device="cuda")
    pd = torch.empty(2 * npix, device="cuda")
    dmt = torch.empty(npix, dtype=torch.int32, device="cuda")
    q._capi.check(whole.ctx.lib.qmri_match_finish_dev(whole.handle, C.c_void_p(xr.data_ptr()), C.c_void_p(xi.data_ptr()), npix, C.c_void_p(merged.data_ptr()),
                                                    C.c_void_p(qmap.data_ptr()), C.c_void_p(pd.data_ptr()), None, C.c_void_p(dmt.data_ptr())))
    whole.ctx.synchronize()
    same = dmt == ref[3]
    assert same.float().mean() > 0.97                         # range splits may resolve fp32 near-ties differently, nothing else
    assert torch.equal(qmap.view(2, npix)[:, same], ref[0][:, same])
    assert int(dmt[0]) == 18
    for h in parts:
        h.close()
    whole.close()


@pytest.mark.parametrize("pipe", ["tensor", "fma"])
def test_matching_is_invariant_to_the_pixel_amplitude(q, monkeypatch, pipe):
    """The reference ranks atoms by abs(ip); the kernels rank by |ip|^2, which halves the fp32 exponent range.  A per-pixel power-of-two
    scale (csrc/match_score.cuh) keeps the squared scores in range: signatures of amplitude 2^-100 ... 2^83 match the same atoms as
    at amplitude 1, and mt / pd follow the original amplitude."""
    import benchdata
    monkeypatch.setenv("QMRI_K2_PIPE", pipe)
    d = benchdata.make_dictionary(K_target=4000, cut=3, seed=4)
    X = _pixels(d, 600, True, seed=9)
    par = {"f": {"qout": 1, "pdout": 1, "dmout": 1, "mtout": 1}}
    ref = q.mrf_dtm_cpu(d, {"X": X}, par)
    for amp in (2.0 ** -100, 2.0 ** -75, 2.0 ** -40, 2.0 ** 40, 2.0 ** 83):   # powers of two: the fp32 signatures are exact multiples
        out = q.mrf_dtm_cpu(d, {"X": X * amp}, par)
        assert np.array_equal(out["dm"], ref["dm"]), (pipe, amp)
        assert np.array_equal(out["qmap"], ref["qmap"])
        assert rel_l2(out["mt"].astype(np.float64), ref["mt"].astype(np.float64) * amp) < 1e-5
        assert rel_l2(out["pd"].astype(np.complex128), ref["pd"].astype(np.complex128) * amp) < 1e-5
