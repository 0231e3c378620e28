cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1"
(timeout 400 $TR --master-port 29551 tests/multigpu_worker.py 2>&1 | grep -E "MULTIGPU|P2P|Error|assert|Traceback" | cut -c1-3500) > gpurun_out/r2_mg8_worker.log
cat gpurun_out/r2_mg8_worker.log
python - <<'PY' 2>&1 | grep -E "REGMG|rror" | cut -c1-300
import os, sys, subprocess
code = r'''
import os, sys, numpy as np, torch, torch.distributed as dist
sys.path.insert(0, "."); sys.path.insert(0, "gpu-accelerated-point-cloud-registration-using-hierarchical-gmm_b200")
import hgmm_b200
from hgmm_b200 import hgmm as H, dist as hd
rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local); dist.init_process_group("nccl", device_id=torch.device("cuda", local))
S = np.load("tests/golden/bun000_xyz.npy"); T = np.load("tests/golden/bun045_xyz.npy")
eng = hgmm_b200.Engine(local); hd.attach_communicator(eng)
init = S[H.reference_init_indices(3)]
eng.set_points(hd.shuffled_shard(S, rank, world, seed=1), total=len(S)); eng.reg_set_target(hd.shuffled_shard(T, rank, world, seed=2))
eng.fit_tree(init, 3, ls=20.0, ld=1e-4, sig2=4e-4, ll_mode="estep", want_current=False, want_outputs=False)
best = 1e9
for _ in range(5):
    dist.barrier()
    rot, t, q, it, _h = eng.register_tree(solver="twist_lstsq", maxiter=20, tol=0.0)
    best = min(best, float(eng.last_timing_ms()[0]))
if rank == 0: print("REGMG world=%d reg_p2p=%s: %d iterations %.3f ms -> %.1f us/iteration" % (world, os.environ.get("HGMM_REG_P2P", "0"), it, best, best * 1e3 / it), flush=True)
eng.comm_destroy(); dist.destroy_process_group()
'''
open("/tmp/regmg.py", "w").write(code)
for v in ("0", "1"):
    env = dict(os.environ, HGMM_REG_P2P=v)
    r = subprocess.run(["timeout", "120", sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "8", "--master-addr", "127.0.0.1", "--master-port", "2956" + v, "/tmp/regmg.py"], env=env, capture_output=True, text=True)
    print(r.stdout[-600:]); print(r.stderr[-300:])
PY
timeout 500 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29571 bench.py --gpus 8 --steps 5 --warmup 3 > gpurun_out/r2_bench_n8.json 2> gpurun_out/r2_bench_n8.err
echo "bench rc=$?"
