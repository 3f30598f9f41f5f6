"""Per-layer device timing of the tensor-mode denoiser (CUDA events around every launch; QMRI_PROFILE=1).
Usage (GPU box): python profiles/layer_profile.py [slices]"""
import os, sys, ctypes as C
os.environ["QMRI_PROFILE"] = "1"
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "qmri-pnp-recon-poc_b200")]
import numpy as np, torch
import qmri_b200 as q
import bench
S = int(sys.argv[1]) if len(sys.argv) > 1 else 8
ctx = q.Context(0)
net = q.UNetRes(bench.make_weights(), in_nc=10, ctx=ctx)
net.set_precision("tc")
x = torch.rand(S * 10 * 224 * 224, device="cuda")
y = torch.empty_like(x)
os.environ.pop("QMRI_PROFILE")
def fwd():
    q._capi.check(ctx.lib.qmri_unetres_forward_dev(net.handle, C.c_void_p(x.data_ptr()), C.c_void_p(y.data_ptr()), None, None, S, 224, 224))
for _ in range(3):
    fwd()
ctx.synchronize()
os.environ["QMRI_PROFILE"] = "1"
fwd()
ctx.synchronize()
