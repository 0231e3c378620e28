"""Probe (not a bench line): configs[4] sharded over the ranks of a torchrun launch; per-rank device time of the tree build.
    torchrun --nproc-per-node N profiles/probe_c5_multi.py [n_points] [L] [reps]        (HGMM_TREE_PROF=1 for the phase clocks)"""
import json, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "gpu-accelerated-point-cloud-registration-using-hierarchical-gmm_b200"))
import numpy as np, torch, torch.distributed as dist, hgmm_b200
from hgmm_b200 import dist as hd, hgmm as H, synth
n = int(sys.argv[1]) if len(sys.argv) > 1 else 1000000
L = int(sys.argv[2]) if len(sys.argv) > 2 else 5
reps = int(sys.argv[3]) if len(sys.argv) > 3 else 4
rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local)
if world > 1:
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
cache = "/tmp/hgmm_lidar_%d_%d.npy" % (n, 2025)
if rank == 0 and not os.path.exists(cache):
    np.save(cache, synth.lidar_sweep(n, seed=2025))
if world > 1:
    dist.barrier()
P = np.load(cache)
init = P[H.reference_init_indices(L)]
eng = hgmm_b200.Engine(local)
if world > 1:
    hd.attach_communicator(eng)
shard = hd.shuffled_shard(P, rank, world, seed=0) if world > 1 else P
eng.set_points(torch.from_numpy(shard).cuda(), total=n)
ms = []
for r in range(reps):
    if world > 1:
        dist.barrier()
    res = eng.fit_tree(init, L, ls=20.0, ld=1e-4, sig2=4.0, ll_mode="estep", want_current=False, want_outputs=False)
    ms.append(float(eng.last_timing_ms()[0]))
its = int(res["iters"].sum())
print("PROBE " + json.dumps({"rank": rank, "world": world, "n": n, "L": L, "p2p": bool(eng.p2p_enabled) if world > 1 else None, "build_ms": min(ms[1:]) if len(ms) > 1 else ms[0],
                             "iters": res["iters"].tolist(), "us_per_iteration": (min(ms[1:]) if len(ms) > 1 else ms[0]) * 1e3 / max(its, 1)}), flush=True)
if world > 1:
    eng.comm_destroy()
    dist.destroy_process_group()
