cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
(timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "registration_steps" 2>&1 | grep -E "v = |passed|failed" | cut -c1-1200) > gpurun_out/r2_dbgA.log
(HGMM_LIB_PATH=$GRAFT_REPO_ROOT/gpu-accelerated-point-cloud-registration-using-hierarchical-gmm_b200/hgmm_b200/libhgmm_vb.so timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q -k "registration_steps or registration_loop or flat_registration" 2>&1 | grep -E "v = |passed|failed|FAILED" | cut -c1-1200) > gpurun_out/r2_dbgB.log
cat gpurun_out/r2_dbgA.log gpurun_out/r2_dbgB.log
