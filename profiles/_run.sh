cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
(timeout 300 $TR --master-port 29551 tests/multigpu_worker.py 2>&1 | grep -E "MULTIGPU|P2P|Error|assert|Traceback" | cut -c1-3500) > gpurun_out/r2_mg2_worker.log; echo "worker rc=${PIPESTATUS[0]}"
cat gpurun_out/r2_mg2_worker.log
timeout 400 $TR --master-port 29571 bench.py --gpus 2 --steps 5 --warmup 3 > gpurun_out/r2_bench_n2.json 2> gpurun_out/r2_bench_n2.err; echo "bench rc=$?"
python - <<'PY'
import json
d = json.loads(open("gpurun_out/r2_bench_n2.json").read().strip().splitlines()[-1])
print("value %.1f e2e %.1f parity ok %s" % (d["value"], d["e2e"]["value"], d["parity_vs_single"]["ok"]), d["parity_vs_single"]["max_rel_fro"])
PY
