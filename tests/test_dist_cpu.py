"""CPU, world_size 2 over gloo: the N>1 host logic -- shard bounds, the seeded shuffle, the unique-id
broadcast channel and the additivity of the packed sufficient statistics the engine all-reduces."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from conftest import ROOT, PKG


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, q):
    import sys
    for p in (ROOT, PKG):
        if p not in sys.path:
            sys.path.insert(0, p)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from hgmm_b200 import dist as hd
    from oracle import hgmm_tree, synth
    X = synth.bunny_like(3000, seed=3).astype(np.float64)
    # every rank derives the same permutation and takes its own contiguous slice
    shard = hd.shuffled_shard(X, rank, world, seed=11)
    lo, hi = hd.shard_bounds(len(X), rank, world)
    assert shard.shape[0] == hi - lo
    # bootstrap channel used for the NCCL unique id
    payload = bytes(range(128)) if rank == 0 else b""
    got = hd.broadcast_bytes(payload, 0)
    assert got == bytes(range(128))
    # bootstrap channel used for the 64-byte IPC handles of the peer-memory exchange windows (rank order preserved)
    hs = hd.allgather_bytes(bytes([rank + 1]) * 64)
    assert hs == [bytes([r + 1]) * 64 for r in range(world)]
    # one tree E-step: moments are additive over shards -> all-reduce == unsharded
    L = 2
    nt = hgmm_tree.n_total(L)
    init = X[hgmm_tree.reference_init_indices(L)]
    pi = np.full(nt, 1 / 8); mu = init.copy(); cov = np.tile(np.eye(3) * 4e-4, (nt, 1, 1))
    parent = -np.ones(len(shard), dtype=np.int64)
    M0, M1, M2, _, den = hgmm_tree.tree_e_step(shard, pi, mu, cov, parent, nt)
    packed = np.concatenate([M0[:8], M1[:8].ravel(), M2[:8].ravel(), [np.log(np.maximum(den, 1e-15)).sum(), len(shard)]])
    tot = hd.allreduce_moments_host(packed)
    if rank == 0:
        q.put(tot)
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_sharded_moments_match_single_rank():
    from oracle import hgmm_tree, synth
    from hgmm_b200 import dist as hd
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    tot = q.get(timeout=120)
    for p in procs:
        p.join(timeout=120)
        assert p.exitcode == 0
    X = synth.bunny_like(3000, seed=3).astype(np.float64)
    L = 2
    nt = hgmm_tree.n_total(L)
    init = X[hgmm_tree.reference_init_indices(L)]
    pi = np.full(nt, 1 / 8); mu = init.copy(); cov = np.tile(np.eye(3) * 4e-4, (nt, 1, 1))
    M0, M1, M2, _, den = hgmm_tree.tree_e_step(X, pi, mu, cov, -np.ones(len(X), dtype=np.int64), nt)
    want = np.concatenate([M0[:8], M1[:8].ravel(), M2[:8].ravel(), [np.log(np.maximum(den, 1e-15)).sum(), len(X)]])
    assert np.allclose(tot, want, rtol=1e-10, atol=1e-12)


def test_shard_bounds_partition_the_cloud():
    from hgmm_b200 import dist as hd
    for n in (0, 1, 7, 40256, 1000000):
        for w in (1, 2, 3, 4, 8):
            edges = [hd.shard_bounds(n, r, w) for r in range(w)]
            assert edges[0][0] == 0 and edges[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(edges, edges[1:]))
            sizes = [b - a for a, b in edges]
            assert max(sizes) - min(sizes) <= 1
    with pytest.raises(ValueError):
        hd.shard_bounds(10, 2, 2)
