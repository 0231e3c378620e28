cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 300 python profiles/probe_flat8_phases.py > gpurun_out/r2_flat8_phases.txt 2>&1
cat gpurun_out/r2_flat8_phases.txt | cut -c1-200
