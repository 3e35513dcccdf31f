mkdir -p gpurun_out
timeout -s KILL 600 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/pytest_gpu.log
python __graft_entry__.py smoke > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?"; tail -2 gpurun_out/smoke.log
python bench.py > gpurun_out/bench_n1.log 2> gpurun_out/bench_n1.err; echo "bench rc=$?"; tail -3 gpurun_out/bench_n1.err
python - <<'PY'
import json
r = json.loads([x for x in open("gpurun_out/bench_n1.log") if x.startswith("{")][-1])
print("value=%.1f G/s ms/step=%.4f roofline=%.3f e2e=%.2f G/s (%.1f ms/step) cpu=%s" % (r["value"]/1e9, r["ms_per_step"], r["roofline"]["frac"], r["e2e"]["value"]/1e9, r["e2e"]["ms_per_step"], r["cpu_baseline"]))
for k, v in r["other_workloads"].items(): print(k, {a: round(b, 2) for a, b in v.items()})
PY
python bench.py --impl reference --steps 5 --warmup 1 > gpurun_out/bench_ref.log 2> gpurun_out/bench_ref.err; echo "ref rc=$?"; cut -c1-300 gpurun_out/bench_ref.log
