"""CPU emulation of the numerics of a tcgen05 kind::tf32 matching kernel (K2 tensor variant), before writing it.

Scheme: every fp32 value a is split a = hi + lo with hi = tf32(a) (11 significant bits), lo = tf32(a - hi); the contraction
<d, x> over C = 10 channels is laid out along a K = 32 axis as [x_hi | x_lo | x_hi | 0 0] . [d_hi | d_hi | d_lo | 0 0]
(x_hi d_hi + x_lo d_hi + x_hi d_lo; the dropped x_lo d_lo term is 2^-22), accumulated in fp32.  Questions answered here:
  1. how far are the tensor scores from the fp32 / fp64 scores (relative to |x|^2)?
  2. is the fp64 winner always inside the tensor top-2 when the fp64 top-2 gap is >= 1e-6 (the parity rule)?
  3. for comparison: single-term TF32 (x_hi d_hi only) - how many atoms would a candidate list need?
"""
import sys
import numpy as np
sys.path.insert(0, ".")
import benchdata


def tf32(a, truncate=False):
    b = np.ascontiguousarray(a, dtype=np.float32).view(np.uint32).astype(np.uint64)
    if truncate:
        b = b & np.uint64(0xFFFFE000)
    else:
        b = (b + np.uint64(0x0FFF) + ((b >> np.uint64(13)) & np.uint64(1))) & np.uint64(0xFFFFE000)
    return b.astype(np.uint32).view(np.float32).reshape(np.shape(a))


def split(a):
    hi = tf32(a)
    lo = tf32((np.asarray(a, np.float32) - hi).astype(np.float32))
    return hi, lo


def scores(D, xr, xi, mode):
    """|<d, x>|^2 for all (atom, pixel); fp32 accumulation emulated by float32 matmul of the exactly representable operands."""
    if mode == "f64":
        return (D.astype(np.float64) @ xr.T.astype(np.float64)) ** 2 + (D.astype(np.float64) @ xi.T.astype(np.float64)) ** 2
    if mode == "f32":
        return (D @ xr.T) ** 2 + (D @ xi.T) ** 2
    dh, dl = split(D)
    out = 0
    for x in (xr, xi):
        xh, xl = split(x)
        if mode == "tf32x1":
            ip = dh @ xh.T
        else:
            ip = (dh @ xh.T + (dh @ xl.T + dl @ xh.T)).astype(np.float32)
        out = out + ip.astype(np.float32) ** 2
    return out.astype(np.float32)


def main():
    K = int(sys.argv[1]) if len(sys.argv) > 1 else 100000
    d = benchdata.make_dictionary(K_target=K, cut=3, seed=0)
    D = d["D"]
    rng = np.random.default_rng(0)
    npix = 2000
    idx = rng.integers(0, D.shape[0], npix)
    pd = rng.uniform(0.3, 1.0, npix)[:, None]
    ph = np.exp(1j * rng.uniform(0, 2 * np.pi, npix))[:, None]
    X = (D[idx] * d["normD"][idx, None] * pd) * ph
    X = X + 0.02 * np.abs(X).max() * (rng.standard_normal(X.shape) + 1j * rng.standard_normal(X.shape))
    xr, xi = X.real.astype(np.float32), X.imag.astype(np.float32)
    n2 = (xr.astype(np.float64) ** 2 + xi.astype(np.float64) ** 2).sum(1)
    s64 = scores(D, xr, xi, "f64")
    order = np.argsort(-s64, axis=0)[:2]
    best, second = s64[order[0], np.arange(npix)], s64[order[1], np.arange(npix)]
    gap = (np.sqrt(best) - np.sqrt(second)) / np.sqrt(best)
    decided = gap >= 1e-6
    print(f"K = {D.shape[0]}, pixels = {npix}, decided (fp64 top-2 gap >= 1e-6): {decided.mean():.3f}; median gap {np.median(gap):.2e}")
    for mode in ("f32", "tf32x3", "tf32x1"):
        s = scores(D, xr, xi, mode).astype(np.float64)
        err = np.abs(s - s64).max(0) / n2
        top = np.argsort(-s, axis=0)[:2]
        win1 = top[0] == order[0]
        in2 = win1 | (top[1] == order[0])
        # candidates a threshold scheme would have to rescore: atoms within 2 * max error of the maximum
        thr = s.max(0) - 2 * err * n2
        ncand = (s >= thr[None, :]).sum(0)
        print(f"{mode:7s}: max |s - s64| / |x|^2 = {err.max():.2e} (median {np.median(err):.2e}); top-1 = fp64 winner on {win1[decided].mean():.4f} of decided "
              f"pixels, fp64 winner inside top-2 on {in2[decided].mean():.4f}; candidates within 2 x error: median {int(np.median(ncand))}, max {ncand.max()}")


if __name__ == "__main__":
    main()
