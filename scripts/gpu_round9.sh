mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
tail -4 gpurun_out/pytest_gpu.log
python scripts/bench_kernels.py --reps 30 --only bandit_rollin > gpurun_out/k1.jsonl 2> gpurun_out/k1.err
python scripts/bench_kernels.py --reps 30 --only darkroom > gpurun_out/k2.jsonl 2> gpurun_out/k2.err
python - <<'PY'
import json
for f in ("gpurun_out/k1.jsonl", "gpurun_out/k2.jsonl"):
    for l in open(f):
        r = json.loads(l)
        print("%-75s %9.4f ms %8.2f Gsteps/s %7.1f GB/s frac %.3f" % (r["kernel"][:75], r["ms_mean"], r["env_steps_per_s"]/1e9, r["achieved_gbs"], r["frac_of_measured_hbm_peak"]))
PY
python bench.py --steps 200 --warmup 10 --no-cpu-baseline > gpurun_out/bench_n1.log 2> gpurun_out/bench_n1.err; cut -c1-330 gpurun_out/bench_n1.log
