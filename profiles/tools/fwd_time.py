"""Times the tensor-mode denoiser forward at S slices (device events, 2 warm-ups, `reps` timed): ms per forward, algorithmic
TFLOP/s.  Usage (GPU box): [QMRI_NET_CHUNK=n] python profiles/tools/fwd_time.py [slices] [reps]"""
import os, sys, ctypes as C
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path[:0] = [ROOT, os.path.join(ROOT, "qmri-pnp-recon-poc_b200")]
import numpy as np, torch
import qmri_b200 as q, bench
S = int(sys.argv[1]) if len(sys.argv) > 1 else 120
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 6
ctx = q.Context(0)
stream = torch.cuda.Stream(); torch.cuda.set_stream(stream); ctx.set_stream(stream.cuda_stream)
net = q.UNetRes(bench.make_weights(), in_nc=10, ctx=ctx)
net.set_precision("tc")
x = torch.rand(S * 10 * 224 * 224, device="cuda")
y = torch.empty_like(x)
def fwd():
    q._capi.check(ctx.lib.qmri_unetres_forward_dev(net.handle, C.c_void_p(x.data_ptr()), C.c_void_p(y.data_ptr()), None, None, S, 224, 224))
for _ in range(2):
    fwd()
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record(stream)
for _ in range(reps):
    fwd()
e1.record(stream); torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / reps
print(f"forward S={S} chunk={os.environ.get('QMRI_NET_CHUNK', 'default')}: {ms:.3f} ms, {ms / S * 15:.3f} ms per 15 slices, {net.flops(S, 224, 224) / ms / 1e9:.1f} TFLOP/s algorithmic, checksum {float(y.double().sum()):.6e}")
