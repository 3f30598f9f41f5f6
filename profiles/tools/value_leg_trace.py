"""Per-step device time of the bench's `value` leg (diagnostic): same setup order as bench.run_ours, 12 steps, one event pair each.
Usage (GPU box): python profiles/tools/value_leg_trace.py [nvml|nonvml] [flush|noflush]"""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path[:0] = [ROOT, os.path.join(ROOT, "qmri-pnp-recon-poc_b200")]
import numpy as np, torch
import qmri_b200 as q, bench
use_nvml = (sys.argv[1] if len(sys.argv) > 1 else "nvml") == "nvml"
use_flush = (sys.argv[2] if len(sys.argv) > 2 else "flush") == "flush"
ctx = q.Context(0)
stream = torch.cuda.Stream(); torch.cuda.set_stream(stream); ctx.set_stream(stream.cuda_stream)
P = q.setup_subsampling_spiralgrided(224, 224, 771, np.eye(10), ctx=ctx)
F = q.fft_operator(P)
X = bench.synthetic_slices(1, 1000)
Y = F.forward(X); X0 = F.adjoint(Y)
net = q.UNetRes(bench.make_weights(), in_nc=10, ctx=ctx)
net.set_precision("tc")
sess = q.AdmmSession({"iter": 100, "gamma": 0.05, "F": F, "X0": X0, "net": net, "denoiser_type": "single_level"}, 1)
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
sess.upload(Y, X0)
if use_nvml:
    s = bench.ClockSampler(0); s.start()
ev = [torch.cuda.Event(enable_timing=True) for _ in range(13)]
torch.cuda.synchronize()
t0 = time.time()
ev[0].record(stream)
for i in range(12):
    if use_flush:
        flush.zero_()
    sess.run(100)
    ev[i + 1].record(stream)
t_enq = time.time() - t0
torch.cuda.synchronize()
print("nvml" if use_nvml else "no nvml", "flush" if use_flush else "no flush", "host enqueue time %.1f ms;" % (t_enq * 1e3),
      "ms per step:", [round(ev[i].elapsed_time(ev[i + 1]), 1) for i in range(12)])
