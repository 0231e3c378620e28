cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 600 python profiles/dump_tree.py persist 2>&1 | grep -v Warning | tail -5
HGMM_TREE_LEGACY=1 timeout 600 python profiles/dump_tree.py legacy 2>&1 | grep -v Warning | tail -5
(timeout 1500 python -m pytest tests -m gpu -q -x --deselect tests/test_gpu_parity.py::test_tree_config_size_matches_oracle_golden --deselect tests/test_gpu_parity.py::test_tree_config_size_long_run_stays_within_the_fp32_storage_envelope --deselect tests/test_gpu_parity.py::test_tree_config_size_converged_against_oracle_golden 2>&1 | tail -30) > gpurun_out/r2_all.log
grep -E "passed|failed|FAILED|Error" gpurun_out/r2_all.log | head
