"""Device time of consecutive qmri_admm_run calls, with and without a host sync between them (diagnostic)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path[:0] = [ROOT, os.path.join(ROOT, "qmri-pnp-recon-poc_b200")]
import numpy as np, torch
import qmri_b200 as q, bench
iters = int(sys.argv[1]) if len(sys.argv) > 1 else 8
ctx = q.Context(0)
stream = torch.cuda.Stream(); torch.cuda.set_stream(stream); ctx.set_stream(stream.cuda_stream)
P = q.setup_subsampling_spiralgrided(224, 224, 771, np.eye(10), ctx=ctx)
F = q.fft_operator(P)
X = bench.synthetic_slices(1, 1)
Y = F.forward(X); X0 = F.adjoint(Y)
net = q.UNetRes(bench.make_weights(), in_nc=10, ctx=ctx)
sess = q.AdmmSession({"iter": iters, "gamma": 0.05, "F": F, "X0": X0, "net": net, "denoiser_type": "single_level"}, 1)
sess.upload(Y, X0)
for sync in (True, False):
    for _ in range(2):
        sess.run(iters)
    torch.cuda.synchronize()
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(7)]
    ev[0].record(stream)
    for i in range(6):
        sess.run(iters)
        ev[i + 1].record(stream)
        if sync:
            torch.cuda.synchronize()
    torch.cuda.synchronize()
    print("sync between runs" if sync else "no sync", [round(ev[i].elapsed_time(ev[i + 1]), 2) for i in range(6)], "ms per run of", iters, "iterations")
