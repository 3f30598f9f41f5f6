"""Times the two K2 pipes (FP32-FMA kernel vs tcgen05 tf32 kernel) on one slice of complex pixels; prints px-atoms/s.
    python profiles/tools/k2_pipes.py [K ...]
Used for the pipe choice in qmri_match_keys_dev (profiles/r02_k2_pipes.md) and as the ncu target (-k regex:match)."""
import ctypes as C
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path[:0] = [ROOT, os.path.join(ROOT, "qmri-pnp-recon-poc_b200")]
import qmri_b200 as q  # noqa: E402

Ks = [int(a) for a in sys.argv[1:]] or [100000]
ctx = q.Context(0)
stream = torch.cuda.Stream()
torch.cuda.set_stream(stream)
ctx.set_stream(stream.cuda_stream)
npix = 224 * 224
reps = int(os.environ.get("K2_REPS", "5"))
for K in Ks:
    rng = np.random.default_rng(0)
    kind = os.environ.get("K2_DICT", "random")
    if kind == "mrf":   # the synthetic FISP-like dictionary of the benchmark: atoms ordered along (T1, T2), scores vary smoothly with the index
        import benchdata
        dd = benchdata.make_dictionary(K_target=K, cut=3, seed=0)
        D = dd["D"]
        K = D.shape[0]
        d = q.Dictionary(dd, ctx=ctx)
        idx = rng.integers(0, K, npix)
        X = D[idx].astype(np.float64) * rng.uniform(0.3, 1.0, (npix, 1)) * np.exp(1j * rng.uniform(0, 2 * np.pi, (npix, 1)))
        X = X + 0.02 * np.abs(X).max() * (rng.standard_normal(X.shape) + 1j * rng.standard_normal(X.shape))
        xr = torch.from_numpy(np.ascontiguousarray(X.real.T.astype(np.float32))).cuda().reshape(-1)
        xi = torch.from_numpy(np.ascontiguousarray(X.imag.T.astype(np.float32))).cuda().reshape(-1)
    else:
        D = rng.standard_normal((K, 10)).astype(np.float32)
        D /= np.linalg.norm(D, axis=1, keepdims=True)
        d = q.Dictionary({"D": D, "normD": np.ones(K, np.float32), "lut": rng.random((K, 2)).astype(np.float32)}, ctx=ctx)
        xr = torch.randn(10 * npix, device="cuda")
        xi = torch.randn(10 * npix, device="cuda")
    keys = {}
    for pipe in os.environ.get("K2_PIPES", "fma,tensor").split(","):
        os.environ["QMRI_K2_PIPE"] = pipe
        k = torch.zeros(npix, dtype=torch.int64, device="cuda")

        def run():
            q._capi.check(ctx.lib.qmri_match_keys_dev(d.handle, C.c_void_p(xr.data_ptr()), C.c_void_p(xi.data_ptr()), npix, C.c_void_p(k.data_ptr())))
        for _ in range(2):
            run()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        for _ in range(reps):
            run()
        e1.record(stream)
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / reps
        keys[pipe] = k.clone()
        print(f"K = {K:8d} ({kind}) pipe {pipe:6s}: {ms:8.3f} ms per slice, {npix * K / (ms * 1e-3):.3e} px-atoms/s, {npix * K / (ms * 1e-3) / (148 * 1.965e9):.2f} scores/clk/SM")
    if len(keys) == 2:
        a, b = keys["fma"], keys["tensor"]
        print(f"           identical keys on {float((a == b).float().mean()):.5f} of the pixels")
    d.close()
