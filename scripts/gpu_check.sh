#!/usr/bin/env bash
# One GPU-box pass: parity tests, smoke, per-kernel table, default bench.  Usage (from the dev container):
#   gpurun --timeout 1500 -- 'bash scripts/gpu_check.sh'
mkdir -p gpurun_out
timeout -s KILL 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/pytest_gpu.log
python __graft_entry__.py smoke > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?"; tail -1 gpurun_out/smoke.log
python scripts/bench_kernels.py --reps 12 > gpurun_out/kernels.jsonl 2> gpurun_out/kernels.err; echo "kernels rc=$?"
python - <<'PY'
import json
for l in open("gpurun_out/kernels.jsonl"):
    r = json.loads(l)
    print("%-78s %9.4f ms %9.3f Gsteps/s %7.1f GB/s frac %.3f" % (r["kernel"][:78], r["ms_mean"], r["env_steps_per_s"]/1e9, r["achieved_gbs"], r["frac_of_measured_hbm_peak"]))
PY
python bench.py > gpurun_out/bench_n1.log 2> gpurun_out/bench_n1.err; echo "bench rc=$?"
python - <<'PY'
import json
r = json.loads([x for x in open("gpurun_out/bench_n1.log") if x.startswith("{")][-1])
print("bench: value=%.1f G env-steps/s ms/step=%.4f roofline=%.3f e2e=%.2f G/s cpu=%.2f M/s (%d cores) clocks=%s" % (
    r["value"]/1e9, r["ms_per_step"], r["roofline"]["frac"], r["e2e"]["value"]/1e9, r["cpu_baseline"]["value"]/1e6, r["cpu_baseline"]["cores"], r["clocks"]))
for k, v in r["other_workloads"].items(): print("  ", k, {a: round(b, 2) for a, b in v.items()})
PY
