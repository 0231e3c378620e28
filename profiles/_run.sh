cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
(timeout 1800 python -m pytest tests -m gpu -q 2>&1 | tail -40) > gpurun_out/r2_all.log
grep -E "passed|failed|FAILED|^E  " gpurun_out/r2_all.log | cut -c1-300 | head -30
