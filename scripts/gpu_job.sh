#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_rollout_gpu.py -m gpu -x -q 2>&1 | tail -3
timeout 600 python scripts/time_dropin.py > gpurun_out/time_dropin.json 2> gpurun_out/time_dropin.err; echo "dropin rc=$?"; cat gpurun_out/time_dropin.json; tail -3 gpurun_out/time_dropin.err
DPT_BENCH_ENVS=1000000 timeout 600 python bench.py --no-cpu-baseline --no-other --no-online-eval --steps 20 --e2e-steps 3 > gpurun_out/bench_1m.json 2> gpurun_out/bench_1m.err; echo "1M rc=$?"
python -c "
import json; l=json.load(open('gpurun_out/bench_1m.json')); print('1M envs: value %.1f G'%(l['value']/1e9), 'ms/step %.3f'%l['ms_per_step'], 'frac %.3f'%l['roofline']['frac'], 'e2e %.2f G'%(l['e2e']['value']/1e9))"
tail -2 gpurun_out/bench_1m.err
