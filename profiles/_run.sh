cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
(timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -15 | cut -c1-600) > gpurun_out/r2_final_tests.log
cat gpurun_out/r2_final_tests.log
timeout 120 python -c "import __graft_entry__ as g; g.smoke(); print('SMOKE OK')" 2>&1 | tail -3
timeout 400 python bench.py > gpurun_out/r2_bench_n1.json 2> gpurun_out/r2_bench_n1.err; echo "bench rc=$?"
python - <<'PY'
import json
d = json.loads(open("gpurun_out/r2_bench_n1.json").read().strip().splitlines()[-1])
print("value %.1f ms/step %.4f e2e %.1f sweep %.2f us frac %.3f launches %s" % (d["value"], d["ms_per_step"], d["e2e"]["value"], d["roofline"]["avg_launch_us"], d["roofline_fp32"]["frac"], d.get("gpu_launches")))
print({k: (v if not isinstance(v, dict) else "...") for k, v in d.items() if k in ("metric", "unit", "n_gpus", "steps", "warmup", "vs_baseline", "dtype")})
PY
