cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
nvidia-smi -L | wc -l
(timeout 400 python -m pytest tests/test_multigpu.py -m gpu -q -k "8 or two_contexts" 2>&1 | tail -15) > gpurun_out/r2_mg8_test.log
grep -E "MULTIGPU|P2P|passed|failed|rror" gpurun_out/r2_mg8_test.log | cut -c1-1500
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29531 bench.py --gpus 8 --steps 10 --warmup 3 > gpurun_out/r2_bench_n8.json 2> gpurun_out/r2_bench_n8.err
echo "bench rc=$?"
tail -c 600 gpurun_out/r2_bench_n8.err
