#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_rollout_gpu.py -m gpu -x -q -k "host" 2>&1 | tail -2
run() { python -c "
import json,sys; l=json.load(open(sys.argv[1])); e=l['e2e']; print(sys.argv[2], 'e2e %.3f G/s'%(e['value']/1e9), 'd2h B/step %.1f'%(e['d2h_bytes_per_step']/62.5e6), 'host peak %.0f'%e.get('host_write_peak_gbs'), 'frac %.3f'%e.get('frac_of_host_peak'))" $1 "$2"; }
for b in 1 2 3 4 6; do
  DPT_HOST_BACKLOG=$b timeout 300 python bench.py --no-cpu-baseline --no-other --no-online-eval --steps 20 > gpurun_out/e2e_b$b.json 2> gpurun_out/e2e_b$b.err; run gpurun_out/e2e_b$b.json "backlog $b"
done
DPT_HOST_COMPACT=1 timeout 300 python bench.py --no-cpu-baseline --no-other --no-online-eval --steps 20 > gpurun_out/e2e_1.json 2>/dev/null; run gpurun_out/e2e_1.json "all compact"
for w in 3 7; do for b in 1 2 4; do
  DPT_HOST_WORKERS=$w DPT_HOST_BACKLOG=$b timeout 300 python bench.py --no-cpu-baseline --no-other --no-online-eval --steps 20 > gpurun_out/e2e_w.json 2>/dev/null; run gpurun_out/e2e_w.json "workers $w backlog $b"
done; done
