mkdir -p gpurun_out
python scripts/sanitize_smoke.py > gpurun_out/sanitize_plain.log 2>&1; echo "smoke rc=$?"
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29532 bench.py --gpus 2 --steps 200 --warmup 10 > gpurun_out/bench_n2.log 2> gpurun_out/bench_n2.err; echo "bench2 rc=$?"
python - <<'PY'
import json
r = json.loads([x for x in open("gpurun_out/bench_n2.log") if x.startswith("{")][-1])
print("N=%d value=%.1f G env-steps/s ms/step=%.4f roofline=%.3f e2e=%.2f G/s stats=%s" % (r["n_gpus"], r["value"]/1e9, r["ms_per_step"], r["roofline"]["frac"], r["e2e"]["value"]/1e9, r["return_stats"]))
PY
