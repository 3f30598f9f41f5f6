"""Multi-GPU plumbing of the path (one process per GPU, ``torch.distributed``).

* Slices are independent (no inter-slice term in ``PnP_ADMM.m`` / ``mrf_dtm_cpu.m``): a slice batch is cut into contiguous
  per-rank ranges and every rank reconstructs its own range - no collective on the reconstruction path.
* Only the atom-sharded dictionary match (BASELINE config 5: a dictionary too large / too slow for one GPU) exchanges
  data: every rank scores ALL pixels against ITS atom range with the fused kernel and produces one packed key per pixel,

      key = float_bits(|<d, x>|^2) << 32 | (0xFFFFFFFF - atom_index)

  whose integer maximum is "largest score, lowest atom index on ties" - MATLAB's first-index rule of
  ``[mt, dm] = max(abs(ip), [], 1)`` (``mrf_dtm_cpu.m:92``).  One max all-reduce over 8 B per pixel (NCCL over
  NVLink on the GPUs, gloo in the CPU tests) gives every rank the global winner; the LUT / normD gather then runs locally.
  Scores are >= 0, so the keys are < 2^63 and order the same as signed int64 - the dtype both backends reduce.

The key arithmetic is restated here in NumPy (``pack_keys`` / ``unpack_keys``) for the host-side tests; the product path
packs keys on the device (csrc/match_kernel.cu).
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from ._capi import check


def slice_shard(n_slices: int, world: int, rank: int) -> tuple[int, int]:
    """Contiguous [begin, end) slice range of ``rank``; the first ``n_slices % world`` ranks take one extra slice."""
    if world < 1 or not (0 <= rank < world):
        raise ValueError(f"bad rank {rank} of {world}")
    base, extra = divmod(int(n_slices), world)
    begin = rank * base + min(rank, extra)
    return begin, begin + base + (1 if rank < extra else 0)


def atom_shard(n_atoms: int, world: int, rank: int) -> tuple[int, int]:
    """Contiguous [begin, end) atom range of ``rank`` (same rule as ``slice_shard``)."""
    return slice_shard(n_atoms, world, rank)


def pack_keys(score_sq, atom_index) -> np.ndarray:
    """(float32 squared score, 0-based atom index) -> uint64 key; NaN scores give key 0 (never win)."""
    s = np.ascontiguousarray(score_sq, dtype=np.float32)
    bits = s.view(np.uint32).astype(np.uint64)
    idx = np.asarray(atom_index, dtype=np.uint64)
    key = (bits << np.uint64(32)) | (np.uint64(0xFFFFFFFF) - idx)
    return np.where(np.isnan(s) | (s < 0), np.uint64(0), key)


def unpack_keys(keys):
    """uint64 keys -> (float32 squared score, 0-based atom index); key 0 (all-NaN pixel) maps to atom 0 like MATLAB's max."""
    k = np.asarray(keys, dtype=np.uint64)
    score = (k >> np.uint64(32)).astype(np.uint32).view(np.float32)
    idx = (np.uint64(0xFFFFFFFF) - (k & np.uint64(0xFFFFFFFF))).astype(np.int64)
    idx = np.where(k == 0, 0, idx)
    return score, idx


def allreduce_keys(keys, group=None):
    """In-place max all-reduce of a torch uint64/int64 key tensor (any device) over ``group``."""
    import torch
    import torch.distributed as dist
    if not dist.is_available() or not dist.is_initialized() or dist.get_world_size(group) == 1:
        return keys
    view = keys.view(torch.int64)  # keys < 2^63: the signed order is the unsigned order
    dist.all_reduce(view, op=dist.ReduceOp.MAX, group=group)
    return keys


def mrf_dtm_sharded(dictionary, x_re, x_im, npix, group=None, want_mt=False, want_dm=True, shared_stream=False):
    """Atom-sharded ``mrf_dtm_cpu`` on device tensors.

    ``dictionary``: a ``Dictionary`` created with ``shard=atom_shard(K, world, rank)``; with ``shard_only=True`` only the shard's
    atoms are resident on this rank (LUT / normD are replicated: 12 B per atom).  ``x_re`` / ``x_im``: planar fp32 torch CUDA
    tensors ``[C][npix]`` (``x_im`` may be None for real data).  Returns torch CUDA tensors ``qmap [Q][npix]``,
    ``pd [npix][2]``, ``mt``, ``dm``.

    Exchange: one max all-reduce of the packed keys (8 B per pixel).  With a shard-only dictionary the owner of each pixel's
    winning atom computes its outputs (zeros elsewhere) and one sum all-reduce over the packed outputs (``(Q + 4)`` x 4 B per
    pixel) gives every rank the result - exact, since every pixel has exactly one non-zero contribution.
    ``shared_stream=True``: the library context launches on torch's current stream (``ctx.set_stream``), so kernels and
    collectives are ordered by the stream and no host synchronisation is needed.
    """
    import torch
    d = dictionary
    dev = x_re.device
    lib = d.ctx.lib
    sync_t = (lambda: None) if shared_stream else (lambda: torch.cuda.synchronize(dev))
    sync_l = (lambda: None) if shared_stream else d.ctx.synchronize
    keys = torch.zeros(npix, dtype=torch.int64, device=dev)
    pim = C.c_void_p(x_im.data_ptr()) if x_im is not None else None
    sync_t()  # x was produced on torch's stream; the library launches on its own unless the stream is shared
    check(lib.qmri_match_keys_dev(d.handle, C.c_void_p(x_re.data_ptr()), pim, npix, C.c_void_p(keys.data_ptr())))
    sync_l()
    allreduce_keys(keys, group)  # 8 B per pixel
    sync_t()
    # one packed output buffer: [qmap Q | pd 2 | mt 1 | dm 1] planes of npix (dm as int32 bits)
    Q = d.Q
    out = torch.empty((Q + 4) * npix, dtype=torch.float32, device=dev)
    qmap, pd = out[:Q * npix], out[Q * npix:(Q + 2) * npix]
    mt, dm = out[(Q + 2) * npix:(Q + 3) * npix], out[(Q + 3) * npix:].view(torch.int32)
    check(lib.qmri_match_finish_dev(d.handle, C.c_void_p(x_re.data_ptr()), pim, npix, C.c_void_p(keys.data_ptr()),
                                    C.c_void_p(qmap.data_ptr()), C.c_void_p(pd.data_ptr()),
                                    C.c_void_p(mt.data_ptr()), C.c_void_p(dm.data_ptr())))
    sync_l()
    if getattr(d, "shard_only", False):
        import torch.distributed as dist
        if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
            # dm travels as int32: summing the owner's index with zeros is exact
            dist.all_reduce(out[:(Q + 3) * npix], op=dist.ReduceOp.SUM, group=group)
            dist.all_reduce(dm, op=dist.ReduceOp.SUM, group=group)
            sync_t()
    return qmap.view(Q, npix), pd.view(npix, 2), (mt if want_mt else None), (dm if want_dm else None)
