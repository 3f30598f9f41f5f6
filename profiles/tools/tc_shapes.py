"""Tensor-mode denoiser vs exact fp32 mode on a matrix of shapes, each in its own process with a timeout
(a protocol bug in a warp-specialised kernel must show up as HANG / ERR per shape, not stall the whole run).
Usage (GPU box): python profiles/tools/tc_shapes.py [S,H,W ...]"""
import os, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
CHILD = r'''
import sys, os
sys.path[:0] = [%r, os.path.join(%r, "qmri-pnp-recon-poc_b200")]
import numpy as np, torch, time
import qmri_b200 as q, bench
S, H, W = map(int, sys.argv[1].split(","))
ctx = q.Context(0)
net = q.UNetRes(bench.make_weights(), in_nc=10, ctx=ctx)
rng = np.random.default_rng(0)
x = rng.random((S, 10, H, W), dtype=np.float32)
net.set_precision("fp32"); ref = net.forward(x)
net.set_precision("tc")
t0 = time.time(); out = net.forward(x); out2 = net.forward(x); dt = time.time() - t0
err = float(np.linalg.norm(out - ref) / np.linalg.norm(ref))
print(f"OK rel={err:.2e} repeat_equal={bool(np.array_equal(out, out2))} t={dt:.2f}s")
''' % (ROOT, ROOT)
shapes = sys.argv[1:] or ["1,224,224", "2,224,224", "8,224,224", "3,56,120", "1,56,120", "3,120,56", "2,64,40", "9,16,24", "1,32,32", "5,112,112"]
for sh in shapes:
    try:
        r = subprocess.run([sys.executable, "-c", CHILD, sh], capture_output=True, text=True, timeout=60)
        if r.returncode == 3:
            print(r.stderr[-12000:])
        msg = r.stdout.strip().splitlines()[-1] if r.stdout.strip() else "ERR " + r.stderr.strip()[-300:]
    except subprocess.TimeoutExpired:
        msg = "HANG (>60 s)"
    print(f"{sh:>12}: {msg}", flush=True)
