"""Multi-GPU plumbing: one process per GPU, points sharded by rank, NCCL communicator bootstrap.

The reference has no distributed code.  The data path needs exactly one
exchange per EM iteration -- an all-reduce of the O(J) sufficient statistics -- which libhgmm issues
itself on its stream (ncclAllReduce, fp64 payload).  torch.distributed is used only as the
bootstrap channel for the 128-byte NCCL unique id (works over gloo or nccl).
"""
import numpy as np


def shard_bounds(n, rank, world):
    """contiguous shard [lo, hi) of n points for `rank` of `world`: sizes differ by <= 1."""
    if world < 1 or not (0 <= rank < world):
        raise ValueError("bad rank/world")
    base, rem = divmod(int(n), int(world))
    lo = rank * base + min(rank, rem)
    hi = lo + base + (1 if rank < rem else 0)
    return lo, hi


def shuffled_shard(points, rank, world, seed=0):
    """the fixed seeded shuffle + contiguous shard of SURVEY.md 8d (C5): every rank computes the same
    permutation and keeps its own slice."""
    pts = np.asarray(points)
    perm = np.random.default_rng(seed).permutation(pts.shape[0])
    lo, hi = shard_bounds(pts.shape[0], rank, world)
    return pts[perm[lo:hi]]


def broadcast_bytes(payload, src=0):
    """broadcast a bytes object over the default torch.distributed group (any backend)."""
    import torch
    import torch.distributed as dist
    dev = "cuda" if dist.get_backend() == "nccl" else "cpu"
    n = len(payload) if dist.get_rank() == src else 0
    ln = torch.tensor([n], dtype=torch.int64, device=dev)
    dist.broadcast(ln, src)
    buf = torch.zeros(int(ln.item()), dtype=torch.uint8, device=dev)
    if dist.get_rank() == src:
        buf.copy_(torch.frombuffer(bytearray(payload), dtype=torch.uint8))
    dist.broadcast(buf, src)
    return bytes(buf.cpu().numpy().tobytes())


def allgather_bytes(payload):
    """every rank's fixed-length bytes object, in rank order, over the default torch.distributed group (any backend)."""
    import torch
    import torch.distributed as dist
    dev = "cuda" if dist.get_backend() == "nccl" else "cpu"
    mine = torch.frombuffer(bytearray(payload), dtype=torch.uint8).to(dev)
    out = [torch.zeros_like(mine) for _ in range(dist.get_world_size())]
    dist.all_gather(out, mine)
    return [bytes(t.cpu().numpy().tobytes()) for t in out]


def attach_communicator(engine, p2p=None):
    """create libhgmm's NCCL communicator for every rank of the default process group and, when the ranks are the GPUs
    of one box (<= 8) that can map each other's memory, the peer-memory exchange windows of the flat fit (NVLink stores
    fused into the M-step kernel instead of ncclAllReduce).  p2p: True / False / None = on unless HGMM_NO_P2P=1.
    Every rank must reach the same decision, so a rank that cannot attach makes all of them fall back to NCCL."""
    import os
    import torch
    import torch.distributed as dist
    from .engine import Engine
    from ._lib import HgmmError
    rank, world = dist.get_rank(), dist.get_world_size()
    uid = Engine.comm_unique_id() if rank == 0 else b""
    uid = broadcast_bytes(uid, 0)
    engine.comm_init(rank, world, uid)
    if p2p is None:
        p2p = os.environ.get("HGMM_NO_P2P", "0") != "1"
    if p2p:
        attach_p2p(engine)
    return rank, world


def attach_p2p(engine):
    """(re)create the peer-memory exchange windows of every rank of the default process group (collective); returns True when
    every rank is attached, else all of them stay on / fall back to NCCL.  Usable again after Engine.p2p_detach()."""
    import torch
    import torch.distributed as dist
    from ._lib import HgmmError
    rank, world = dist.get_rank(), dist.get_world_size()
    ok = False
    if 2 <= world <= 8:
        try:
            handle = engine.p2p_export()
        except HgmmError:
            handle = b"\0" * 64
        handles = allgather_bytes(handle)
        ok = all(any(h) for h in handles)
        if ok:
            try:
                engine.p2p_attach(handles)
            except HgmmError:
                ok = False
        dev = "cuda" if dist.get_backend() == "nccl" else "cpu"
        flag = torch.tensor([1 if ok else 0], dtype=torch.int32, device=dev)
        dist.all_reduce(flag, op=dist.ReduceOp.MIN)
        ok = int(flag.item()) == 1
        if ok:
            ok = _p2p_self_test(engine, rank, world, dev)
        if not ok and engine.p2p_enabled:
            engine.p2p_detach()
    return ok


reattach_p2p = attach_p2p


def _p2p_self_test(engine, rank, world, dev):
    """two EM iterations of a small mixture through the peer-memory exchange on every rank: all ranks must succeed and hold
    bit-identical replicas, otherwise all of them fall back to NCCL (collective decision: every rank reaches both
    all_reduce calls whatever happened locally)."""
    import torch
    import torch.distributed as dist
    from ._lib import HgmmError
    rng = np.random.default_rng(1234 + rank)
    pts = rng.normal(0.0, 1.0, (4096, 3)).astype(np.float32)
    J = 64
    mu0 = np.random.default_rng(99).normal(0.0, 1.0, (J, 3)).astype(np.float32)     # identical on every rank
    cov0 = np.tile(np.eye(3, dtype=np.float32), (J, 1, 1))
    w0 = np.full(J, 1.0 / J, np.float32)
    good, digest = 1, np.zeros(4, np.float64)
    try:
        engine.declare_total_points(0)       # the toy cloud's total comes from the all-reduce, whatever was declared before
        engine.set_points(pts)
        r = engine.fit_flat(mu0, cov0, w0, cov_type="full", max_iter=2)
        digest = np.array([r["means"].astype(np.float64).sum(), r["covs"].astype(np.float64).sum(),
                           r["weights"].astype(np.float64).sum(), float(r["ll"][-1])])
        if not np.isfinite(digest).all():
            good = 0
    except HgmmError:
        good = 0
    lo = torch.from_numpy(np.concatenate([[float(good)], digest])).to(dev)
    hi = lo.clone()
    dist.all_reduce(lo, op=dist.ReduceOp.MIN)
    dist.all_reduce(hi, op=dist.ReduceOp.MAX)
    return bool(lo[0].item() == 1.0 and bool((lo[1:] == hi[1:]).all().item()))


def allreduce_moments_host(local_moments):
    """host-side mirror of the engine's exchange step (sum of packed sufficient statistics over ranks);
    used by the CPU (gloo) tests of the sharding logic and by tooling."""
    import torch
    import torch.distributed as dist
    t = torch.from_numpy(np.ascontiguousarray(local_moments, dtype=np.float64).copy())
    dist.all_reduce(t, op=dist.ReduceOp.SUM)
    return t.numpy()
