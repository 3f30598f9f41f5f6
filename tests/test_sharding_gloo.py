"""N > 1 host logic on CPU: world_size-2 gloo process groups (no GPU).

* slice sharding: contiguous, disjoint, complete; per-rank results concatenate to the unsharded result
* atom-sharded matching (BASELINE config 5): every rank scores its atom range, packs (score, index) keys, one max
  all-reduce -> identical to the unsharded float32 argmax of the oracle, including first-index ties across ranks.
The scoring here is the NumPy float32 mirror of mrf_dtm_cpu.m:91-92 (test infrastructure); on the GPU the same keys come
from csrc/match_kernel.cu and the same all-reduce runs over NCCL (tests/test_gpu_parity.py covers the device side).
"""
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for _p in (ROOT, os.path.join(ROOT, "qmri-pnp-recon-poc_b200")):
    if _p not in sys.path:
        sys.path.insert(0, _p)


def _problem(K=3001, npix=777, C=10, seed=0):
    rng = np.random.default_rng(seed)
    D = rng.standard_normal((K, C)).astype(np.float32)
    D /= np.linalg.norm(D, axis=1, keepdims=True)
    D[1234] = D[17]          # exact duplicate atoms on different ranks: the lower index must win
    D[2999] = D[17]
    x = (rng.standard_normal((npix, C)) + 1j * rng.standard_normal((npix, C))).astype(np.complex64)
    x[5] = 3.0 * D[17]       # a pixel that hits the duplicated atom exactly
    return D, x


def _local_keys(D, x, a0, a1):
    from qmri_b200.sharding import pack_keys
    ip = D[a0:a1] @ x.conj().T                       # ip = dict.D * ctranspose(x), mrf_dtm_cpu.m:91
    sc = (ip.real.astype(np.float32) ** 2 + ip.imag.astype(np.float32) ** 2).astype(np.float32)
    best = sc.argmax(axis=0)                          # first index wins ties, like MATLAB's max
    return pack_keys(sc[best, np.arange(x.shape[0])], best + a0)


def _worker(rank, world, port, ret):
    import torch
    import torch.distributed as dist
    from qmri_b200.sharding import allreduce_keys, atom_shard, slice_shard, unpack_keys
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        D, x = _problem()
        a0, a1 = atom_shard(D.shape[0], world, rank)
        keys = torch.from_numpy(_local_keys(D, x, a0, a1).view(np.int64).copy())
        allreduce_keys(keys)
        score, idx = unpack_keys(keys.numpy().view(np.uint64))
        # slice sharding: every rank reduces its own slices, results gathered in rank order
        S = 15
        s0, s1 = slice_shard(S, world, rank)
        mine = torch.arange(s0, s1, dtype=torch.float64) ** 2        # stand-in for per-slice reconstructions
        parts = [None] * world                                        # ragged shards (8 + 7 slices): gather as objects
        dist.all_gather_object(parts, mine.numpy())
        ret[rank] = (idx.copy(), score.copy(), np.concatenate(parts))
    finally:
        dist.destroy_process_group()


def test_slice_and_atom_shards_are_a_partition():
    from qmri_b200.sharding import atom_shard, slice_shard
    for n, w in [(15, 8), (120, 8), (1, 4), (7, 2), (1048576, 8), (0, 3)]:
        edges = [slice_shard(n, w, r) for r in range(w)]
        assert edges[0][0] == 0 and edges[-1][1] == n
        assert all(edges[r][1] == edges[r + 1][0] for r in range(w - 1))
        sizes = [b - a for a, b in edges]
        assert max(sizes) - min(sizes) <= 1
        assert atom_shard(n, w, 0) == edges[0]
    with pytest.raises(ValueError):
        slice_shard(4, 2, 2)


def test_key_packing_orders_like_matlab_max():
    from qmri_b200.sharding import pack_keys, unpack_keys
    k = pack_keys(np.array([1.0, 1.0, 2.0, np.nan, 0.0], np.float32), np.array([7, 3, 9, 1, 0]))
    assert k[1] > k[0]            # equal score: lower index wins
    assert k[2] > k[1]            # larger score wins
    assert k[3] == 0              # NaN never wins
    assert k[4] > 0               # a zero score still beats "nothing"
    s, i = unpack_keys(k)
    assert i.tolist() == [7, 3, 9, 0, 0] and s[2] == 2.0
    assert np.all(k < np.uint64(1) << np.uint64(63))  # signed max == unsigned max


def test_world2_gloo_atom_sharded_match_equals_unsharded():
    import torch.multiprocessing as mp
    world, port = 2, 29500 + (os.getpid() % 2000)
    with mp.Manager() as mgr:
        ret = mgr.dict()
        mp.spawn(_worker, args=(world, port, ret), nprocs=world, join=True)
        D, x = _problem()
        full_keys = _local_keys(D, x, 0, D.shape[0])
        from qmri_b200.sharding import unpack_keys
        score, idx = unpack_keys(full_keys)
        for r in range(world):
            got_idx, got_score, slices = ret[r]
            assert np.array_equal(got_idx, idx)          # identical indices on every rank, ties included
            assert np.array_equal(got_score, score)
            assert np.array_equal(slices, np.arange(15, dtype=np.float64) ** 2)
        assert idx[5] == 17                               # duplicates at 1234 (rank 0) and 2999 (rank 1) lose to index 17
