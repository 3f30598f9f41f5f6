"""BASELINE configs[4]: dictionary matching against a large synthetic dictionary, atoms sharded over the ranks, one NCCL max
all-reduce of 8 B per pixel.  Checks the sharded result against the unsharded one on rank 0 and prints px*atoms/s.
Usage (GPU box, N GPUs): python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port 29531 \
    profiles/tools/sharded_match_nccl.py [atoms]"""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path[:0] = [ROOT, os.path.join(ROOT, "qmri-pnp-recon-poc_b200")]
import numpy as np, torch, torch.distributed as dist
import qmri_b200 as q
from qmri_b200.sharding import atom_shard, mrf_dtm_sharded

rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local)
if world > 1:
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
K = int(sys.argv[1]) if len(sys.argv) > 1 else 1 << 20
C_CH, npix = 10, 224 * 224
rng = np.random.default_rng(3)                      # same dictionary and data on every rank
D = rng.standard_normal((K, C_CH)).astype(np.float32)
D /= np.linalg.norm(D, axis=1, keepdims=True)
dd = {"D": D, "normD": (1 + rng.random(K)).astype(np.float32), "lut": rng.random((K, 2)).astype(np.float32)}
ctx = q.Context(local)
xr = torch.from_numpy(rng.standard_normal((C_CH, npix)).astype(np.float32)).cuda()
xi = torch.from_numpy(rng.standard_normal((C_CH, npix)).astype(np.float32)).cuda()
d_sh = q.Dictionary(dd, ctx=ctx, shard=atom_shard(K, world, rank))
for _ in range(2):
    out = mrf_dtm_sharded(d_sh, xr, xi, npix, want_mt=True)
torch.cuda.synchronize()
if world > 1:
    dist.barrier()
t0 = time.time()
reps = 5
for _ in range(reps):
    qmap, pd, mt, dm = mrf_dtm_sharded(d_sh, xr, xi, npix, want_mt=True)
torch.cuda.synchronize()
if world > 1:
    dist.barrier()
dt = (time.time() - t0) / reps
if rank == 0:
    d_full = q.Dictionary(dd, ctx=ctx)
    if True:
        # unsharded reference on this rank alone: a dictionary holding every atom, no reduction
        import ctypes as C
        keys = torch.zeros(npix, dtype=torch.int64, device="cuda")
        q._capi.check(ctx.lib.qmri_match_keys_dev(d_full.handle, C.c_void_p(xr.data_ptr()), C.c_void_p(xi.data_ptr()), npix, C.c_void_p(keys.data_ptr())))
        dm0 = torch.empty(npix, dtype=torch.int32, device="cuda")
        qm0 = torch.empty(2 * npix, dtype=torch.float32, device="cuda")
        pd0 = torch.empty(2 * npix, dtype=torch.float32, device="cuda")
        q._capi.check(ctx.lib.qmri_match_finish_dev(d_full.handle, C.c_void_p(xr.data_ptr()), C.c_void_p(xi.data_ptr()), npix, C.c_void_p(keys.data_ptr()),
                                                    C.c_void_p(qm0.data_ptr()), C.c_void_p(pd0.data_ptr()), None, C.c_void_p(dm0.data_ptr())))
        ctx.synchronize()
        qm0 = qm0.view(2, npix)
    same = bool(torch.equal(dm.cpu(), dm0.cpu())) and bool(torch.equal(qmap.cpu(), qm0.cpu()))
    print(f"atom-sharded match: {world} GPU(s), {K} atoms, {npix} complex pixels: {dt * 1e3:.2f} ms per slice (host-timed, incl. the all-reduce) = "
          f"{npix * K / dt / 1e12:.3f}e12 px*atoms/s; equal to the unsharded result: {same}")
    assert same
if world > 1:
    dist.destroy_process_group()
