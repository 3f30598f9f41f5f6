"""Small driver for compute-sanitizer runs over the x-update kernels (streaming, cluster, general V).
Usage (GPU box): compute-sanitizer --tool memcheck|racecheck python profiles/tools/sanitize_k1.py"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path[:0] = [ROOT, os.path.join(ROOT, "qmri-pnp-recon-poc_b200")]
import numpy as np
import qmri_b200 as q
rng = np.random.default_rng(0)
def run(P, C, S):
    F = q.fft_operator(P)
    x = rng.standard_normal((224, 224, C, S)) + 1j * rng.standard_normal((224, 224, C, S))
    y = F.forward(x)
    xa = F.adjoint(y)
    xs = F.xupdate(y, x.real, 0.1 * x, 0.05)
    xk = q.PnP_ADMM(y, {"iter": 3, "gamma": 0.05, "F": F, "X0": xa, "net": lambda v: np.asarray(v)[:, :, :C] * 0.9, "denoiser_type": "single_level"})
    return float(np.abs(xs).sum() + np.abs(xk).sum())
os.environ["QMRI_K1_KERNEL"] = "stream"
print("stream spiral", run(q.setup_subsampling_spiralgrided(224, 224, 771, np.eye(10)), 10, 2))
print("stream epi", run(q.setup_subsampling_epi(224, 224, 1 / 65, np.eye(10)), 10, 1))
os.environ["QMRI_K1_KERNEL"] = "cluster"
print("cluster spiral", run(q.setup_subsampling_spiralgrided(224, 224, 771, np.eye(10)), 10, 1))
os.environ.pop("QMRI_K1_KERNEL")
V = np.linalg.qr(rng.standard_normal((24, 10)))[0]
print("general 24x10", run(q.setup_subsampling_spiralgrided(224, 224, 771, V), 10, 1))
