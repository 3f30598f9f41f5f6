"""CPU emulation of the K1 kernel (same phase functions / tables as the CUDA kernel, thread by
thread) against the float64 oracle.  This is how the kernel arithmetic is checked in a container
without a GPU; the GPU parity tests repeat the comparison on the real kernel."""
import ctypes
import os

import numpy as np
import pytest

from conftest import PKG, rel_l2
from oracle import sampling
from oracle.xupdate import xupdate_exact

LIB = os.path.join(PKG, "lib", "libk1emu.so")
fp = ctypes.POINTER(ctypes.c_float)


def P(a):
    return a.ctypes.data_as(fp) if a is not None else None


@pytest.fixture(scope="module")
def emu():
    if not os.path.exists(LIB):
        pytest.skip("libk1emu.so not built (run __graft_entry__.build())")
    lib = ctypes.CDLL(LIB)
    lib.k1emu_masks.restype = ctypes.c_int
    lib.k1emu_masks.argtypes = [ctypes.c_int, ctypes.c_int, ctypes.c_double, ctypes.c_int, ctypes.c_void_p, ctypes.c_void_p]
    lib.k1emu_run.restype = ctypes.c_int
    lib.k1emu_run.argtypes = [ctypes.c_int] * 3 + [ctypes.c_double, ctypes.c_int, ctypes.c_int] + [fp] * 4 + [ctypes.c_float] + [fp] * 4
    return lib


def planar(a):
    a = np.transpose(a, (2, 1, 0))
    return np.ascontiguousarray(a.real.astype(np.float32)), np.ascontiguousarray(a.imag.astype(np.float32))


def unplanar(re, im):
    return np.transpose(re.astype(np.float64) + 1j * im.astype(np.float64), (2, 1, 0))


@pytest.mark.parametrize("pat,arg", [(0, 771.0), (1, 1 / 65)])
def test_cxx_mask_constructors_equal_oracle(emu, pat, arg):
    for L in (10, 50):
        Po = (sampling.setup_subsampling_spiralgrided(224, 224, 771, np.eye(L)) if pat == 0
              else sampling.setup_subsampling_epi(224, 224, 1 / 65, np.eye(L)))
        n = emu.k1emu_masks(pat, 224, arg, L, None, None)
        idx = np.zeros(n, np.int32)
        fptr = np.zeros(L + 1, np.int64)
        emu.k1emu_masks(pat, 224, arg, L, idx.ctypes.data, fptr.ctypes.data)
        assert np.array_equal(idx, Po.idx) and np.array_equal(fptr, Po.frame_ptr)


@pytest.mark.parametrize("pat,arg,mc", [(0, 771.0, 28), (0, 771.0, 56), (1, 1 / 65, 28), (0, 771.0, 0), (1, 1 / 65, 0)])  # mc 0 = streaming kernel
def test_k1_arithmetic_all_modes(emu, pat, arg, mc, monkeypatch):
    if mc == 0 and pat == 0:
        monkeypatch.setenv("QMRI_K1_QMIN", "1")   # smallest feasible chunk (5 samples): many overflow partials; the default 8 runs for EPI
    C = 3
    Po = (sampling.setup_subsampling_spiralgrided(224, 224, 771, np.eye(C)) if pat == 0
          else sampling.setup_subsampling_epi(224, 224, 1 / 65, np.eye(C)))
    F = sampling.FOperator(Po)
    n = Po.nmeas
    rng = np.random.default_rng(0)
    z = rng.standard_normal((224, 224, C)) + 1j * rng.standard_normal((224, 224, C))
    zr, zi = planar(z)
    yout = np.zeros((n, 2), np.float32)
    emu.k1emu_run(2, mc, pat, arg, C, 1, P(zr), P(zi), None, None, 0.05, None, None, P(yout), None)
    assert rel_l2(yout[:, 0] + 1j * yout[:, 1], F.forward(z)) < 1e-6
    y = rng.standard_normal(n) + 1j * rng.standard_normal(n)
    y32 = np.stack([y.real, y.imag], 1).astype(np.float32)
    ore, oim = np.zeros_like(zr), np.zeros_like(zr)
    emu.k1emu_run(3, mc, pat, arg, C, 1, None, None, None, P(y32), 0.05, P(ore), P(oim), None, None)
    assert rel_l2(unplanar(ore, oim), F.adjoint(y)) < 1e-6
    mm = np.zeros(2, np.float32)
    emu.k1emu_run(1, mc, pat, arg, C, 1, P(zr), P(zi), None, P(y32), 0.05, P(ore), P(oim), None, P(mm))
    xe = xupdate_exact(F, y, z, 0.05)
    assert rel_l2(unplanar(ore, oim), xe) < 1e-6
    assert abs(mm[0] - xe.real.min()) < 1e-5 and abs(mm[1] - xe.real.max()) < 1e-5
    v = rng.standard_normal((224, 224, C))
    vr, _ = planar(v + 0j)
    emu.k1emu_run(0, mc, pat, arg, C, 1, P(zr), P(zi), P(vr), P(y32), 0.05, P(ore), P(oim), None, P(mm))
    x = xupdate_exact(F, y, 2 * v - z, 0.05)
    assert rel_l2(unplanar(ore, oim), x + (z - v)) < 1e-6   # w' = x + u, u = w - v


@pytest.mark.parametrize("pat,arg,L,C", [(0, 771.0, 12, 10), (0, 771.0, 200, 10), (1, 1 / 65, 5, 4)])
def test_general_V_host_tables(emu, pat, arg, L, C):
    """Union of the masks, membership lists, part tiling and (G_u + rho I)^-1 of the general-V path against NumPy."""
    rng = np.random.default_rng(L)
    V = np.asfortranarray(rng.standard_normal((L, C)))
    Po = (sampling.setup_subsampling_spiralgrided(224, 224, 771, V) if pat == 0 else sampling.setup_subsampling_epi(224, 224, 1 / 65, V))
    U = np.unique(np.concatenate([np.asarray(f) for f in Po.frames]))
    emu.k1emu_general.restype = ctypes.c_int
    emu.k1emu_general.argtypes = [ctypes.c_int, ctypes.c_double, ctypes.c_int, ctypes.c_int, ctypes.c_void_p, ctypes.c_double,
                                  ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p]
    ul = np.zeros(len(U), np.int32)
    npart = ctypes.c_int(0)
    Minv = np.zeros((len(U), C, C), np.float32)
    rho = 0.05
    n = emu.k1emu_general(pat, arg, L, C, V.ctypes.data, rho, ul.ctypes.data, ctypes.byref(npart), Minv.ctypes.data)
    assert n == len(U) and np.array_equal(ul, U)
    assert npart.value == -(-len(U) // 3400) or npart.value > len(U) // 3400
    member = {int(k): [] for k in U}
    for i, f in enumerate(Po.frames):
        for k in f:
            member[int(k)].append(i)
    for u in rng.choice(len(U), size=40, replace=False):
        G = sum(np.outer(V[i], V[i]) for i in member[int(U[u])])
        assert np.allclose(Minv[u], np.linalg.inv(G + rho * np.eye(C)), rtol=2e-5, atol=1e-6)


@pytest.mark.parametrize("pat,arg,qmin", [(0, 771.0, "1"), (0, 771.0, "8"), (1, 1 / 65, "8")])
def test_k1_real_image_kernels(emu, pat, arg, qmin, monkeypatch):
    """The real-image streaming kernels of the loop (csrc/xupdate_real.cu: two real columns per complex FFT, folded k-space rows,
    Hermitian-packed inverse) thread by thread on the CPU: m = A v and v + Re(A^H c) against the float64 oracle."""
    monkeypatch.setenv("QMRI_K1_QMIN", qmin)
    emu.k1emu_run_real.restype = ctypes.c_int
    emu.k1emu_run_real.argtypes = [ctypes.c_int, ctypes.c_int, ctypes.c_double, ctypes.c_int, ctypes.c_int] + [fp] * 5 + [ctypes.POINTER(ctypes.c_int)]
    C = 3
    Po = (sampling.setup_subsampling_spiralgrided(224, 224, 771, np.eye(C)) if pat == 0
          else sampling.setup_subsampling_epi(224, 224, 1 / 65, np.eye(C)))
    F = sampling.FOperator(Po)
    n = Po.nmeas
    rng = np.random.default_rng(3)
    v = rng.standard_normal((224, 224, C))
    vr, _ = planar(v + 0j)
    yout = np.zeros((n, 2), np.float32)
    novf = ctypes.c_int(-1)
    assert emu.k1emu_run_real(0, pat, arg, C, 1, P(vr), None, None, P(yout), None, ctypes.byref(novf)) == n
    assert 0 <= novf.value <= 96
    assert rel_l2(yout[:, 0] + 1j * yout[:, 1], F.forward(v)) < 1e-6
    c = rng.standard_normal(n) + 1j * rng.standard_normal(n)
    c32 = np.stack([c.real, c.imag], 1).astype(np.float32)
    out = np.zeros_like(vr)
    mm = np.zeros(2, np.float32)
    emu.k1emu_run_real(1, pat, arg, C, 1, P(vr), P(c32), P(out), None, P(mm), None)
    corr = np.real(F.adjoint(c))
    got = np.transpose(out.astype(np.float64), (2, 1, 0))
    assert rel_l2(got - v, corr) < 2e-6          # the correction itself
    assert rel_l2(got, v + corr) < 1e-6
    assert abs(mm[0] - (v + corr).min()) < 1e-5 and abs(mm[1] - (v + corr).max()) < 1e-5
