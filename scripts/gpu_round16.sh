mkdir -p gpurun_out
N=${1:-2}
timeout -s KILL 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29541 scripts/multi_gpu_check.py > gpurun_out/multi_gpu_check.log 2>&1; echo "mgpu rc=$?"; tail -3 gpurun_out/multi_gpu_check.log
for mode in p2p nccl; do
DPT_BENCH_GATHER=$mode timeout -s KILL 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29542 bench.py --gpus $N --steps 200 --warmup 10 > gpurun_out/bench_$mode.log 2> gpurun_out/bench_$mode.err; echo "bench $mode rc=$?"
python - <<PY
import json
try:
    r = json.loads([x for x in open("gpurun_out/bench_$mode.log") if x.startswith("{")][-1])
    print("$mode N=%d value=%.1f G env-steps/s ms/step=%.4f roofline=%.3f stats=%s | %s" % (r["n_gpus"], r["value"]/1e9, r["ms_per_step"], r["roofline"]["frac"], r["return_stats"], r["config"]["parallelism"][:110]))
except Exception as e:
    print("no json", e); print(open("gpurun_out/bench_$mode.err").read()[-1500:])
PY
done
