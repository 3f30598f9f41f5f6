"""One tensor-mode denoiser forward (after warm-up) for ncu captures.  Usage (GPU box): python profiles/tools/fwd_only.py [slices]"""
import os, sys, ctypes as C
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path[:0] = [ROOT, os.path.join(ROOT, "qmri-pnp-recon-poc_b200")]
import numpy as np, torch
import qmri_b200 as q, bench
S = int(sys.argv[1]) if len(sys.argv) > 1 else 1
ctx = q.Context(0)
net = q.UNetRes(bench.make_weights(), in_nc=10, ctx=ctx)
net.set_precision("tc")
x = torch.rand(S * 10 * 224 * 224, device="cuda")
y = torch.empty_like(x)
for _ in range(3):
    q._capi.check(ctx.lib.qmri_unetres_forward_dev(net.handle, C.c_void_p(x.data_ptr()), C.c_void_p(y.data_ptr()), None, None, S, 224, 224))
ctx.synchronize()
