cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1"
(timeout 400 $TR --master-port 29541 tests/multigpu_worker.py 2>&1 | grep -E "MULTIGPU|P2P|Error|assert" | cut -c1-3000) > gpurun_out/r2_mg8_worker.log
cat gpurun_out/r2_mg8_worker.log
timeout 600 $TR --master-port 29542 bench.py --gpus 8 --steps 10 --warmup 3 > gpurun_out/r2_bench_n8.json 2> gpurun_out/r2_bench_n8.err
echo "bench rc=$?"
(timeout 200 python -m pytest tests/test_multigpu.py -m gpu -q -k "two_contexts" 2>&1 | tail -3)
