mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
tail -5 gpurun_out/pytest_gpu.log
python scripts/bench_kernels.py --reps 8 --only gpt2 > gpurun_out/kernels_gpt2.jsonl 2> gpurun_out/kernels_gpt2.err; echo "rc=$?"
python - <<'PY'
import json
for l in open("gpurun_out/kernels_gpt2.jsonl"):
    r = json.loads(l)
    print("%-75s %9.3f ms %8.4f Gsteps/s %7.1f GB/s frac %.3f" % (r["kernel"][:75], r["ms_mean"], r["env_steps_per_s"]/1e9, r["achieved_gbs"], r["frac_of_measured_hbm_peak"]), r.get("trajs_per_s", ""))
PY
