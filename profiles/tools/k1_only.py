"""K1 (fused x-update) alone at a slice batch larger than L2: device time per launch, for ncu captures.
Usage (GPU box): python profiles/tools/k1_only.py [slices] [reps]"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path[:0] = [ROOT, os.path.join(ROOT, "qmri-pnp-recon-poc_b200")]
import numpy as np, torch
import qmri_b200 as q, bench
S = int(sys.argv[1]) if len(sys.argv) > 1 else 120
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 20
ctx = q.Context(0)
stream = torch.cuda.Stream(); torch.cuda.set_stream(stream); ctx.set_stream(stream.cuda_stream)
P = q.setup_subsampling_spiralgrided(224, 224, 771, np.eye(10), ctx=ctx)
F = q.fft_operator(P)
X = bench.synthetic_slices(S, 5)
Y = F.forward(X)
sess = q.AdmmSession({"iter": 2, "gamma": 0.05, "F": F, "X0": np.zeros((224, 224, 10, S)), "net": lambda v: v}, S)
sess.upload(Y, F.adjoint(Y))
sess.xupdate_only(3)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record(stream); sess.xupdate_only(reps); e1.record(stream); torch.cuda.synchronize()
t = e0.elapsed_time(e1) / reps * 1e-3
bpp = float(sess.xupdate_bytes())   # algorithmic bytes per pixel-channel of the formulation that ran (8: real state, 20: complex state)
state = "real" if bpp == 8 else "complex"
print(f"K1 S={S} state={state} group={os.environ.get('QMRI_K1R_GROUP', 'default')} G={os.environ.get('QMRI_K1R_G', 'auto')}: "
      f"{t*1e6/S:.3f} us per slice-iteration, {bpp*224*224*10*S/t/1e9:.1f} GB/s algorithmic ({bpp:.0f} B/px-ch)")
