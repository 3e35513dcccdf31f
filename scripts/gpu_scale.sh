mkdir -p gpurun_out
N=${1:-8}
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29521 scripts/multi_gpu_check.py > gpurun_out/multi_gpu_check_$N.log 2>&1; echo "mgpu rc=$?"; tail -1 gpurun_out/multi_gpu_check_$N.log
for n in 1 2 4 8; do
  if [ $n -le $N ]; then
    if [ $n -eq 1 ]; then
      python bench.py --gpus 1 --steps 200 --warmup 10 --no-cpu-baseline --no-other > gpurun_out/scale_n$n.log 2> gpurun_out/scale_n$n.err
    else
      python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $((29530+n)) bench.py --gpus $n --steps 200 --warmup 10 > gpurun_out/scale_n$n.log 2> gpurun_out/scale_n$n.err
    fi
    echo "bench n=$n rc=$?"
    python - <<PY
import json
l = [x for x in open("gpurun_out/scale_n$n.log") if x.startswith("{")][-1]
r = json.loads(l)
print("N=%d value=%.1f G env-steps/s ms/step=%.4f roofline=%.3f e2e=%.2f G/s clocks=%s" % (r["n_gpus"], r["value"]/1e9, r["ms_per_step"], r["roofline"]["frac"], r["e2e"]["value"]/1e9, r["clocks"]))
PY
  fi
done
