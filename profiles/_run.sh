cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
(timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -q -k "config_size" 2>&1 | tail -60) > gpurun_out/r2_t3.log
grep -E "passed|failed|FAILED" gpurun_out/r2_t3.log
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
(HGMM_TREE_PROF=1 timeout 600 $TR --master-port 29521 profiles/probe_c5_multi.py 1000000 5 4 2>&1 | grep -E "PROBE|PROF|rror") > gpurun_out/r2_c5_n2.log
(timeout 600 $TR --master-port 29522 profiles/probe_c5_multi.py 1000000 5 4 2>&1 | grep -E "PROBE|PROF|rror") >> gpurun_out/r2_c5_n2.log
(timeout 300 python profiles/probe_c5_multi.py 1000000 5 4 2>&1 | grep -E "PROBE|PROF|rror") >> gpurun_out/r2_c5_n2.log
(HGMM_NO_P2P=1 timeout 600 $TR --master-port 29523 profiles/probe_c5_multi.py 1000000 5 3 2>&1 | grep -E "PROBE|PROF|rror") >> gpurun_out/r2_c5_n2.log
cat gpurun_out/r2_c5_n2.log
