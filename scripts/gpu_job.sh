#!/bin/bash
mkdir -p gpurun_out
N=$(nvidia-smi -L | wc -l)
echo "gpus: $N  cores: $(nproc)"
timeout 900 python -m pytest tests/test_multi_gpu.py tests/test_rollout_gpu.py -m gpu -x -q 2>&1 | tail -3
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus $N --steps 50 --warmup 5 > gpurun_out/bench_n${N}.json 2> gpurun_out/bench_n${N}.err; echo "bench rc=$?"; tail -2 gpurun_out/bench_n${N}.err | cut -c1-300
python -c "
import json,sys; l=json.loads([x for x in open('gpurun_out/bench_n${N}.json') if x.startswith('{')][-1])
print('N=$N: value %.1f G frac %.3f e2e %.2f G (host peak %.0f frac %.2f) d2h/step %.1f'%(l['value']/1e9, l['roofline']['frac'], l['e2e']['value']/1e9, l['e2e'].get('host_write_peak_gbs') or 0, l['e2e'].get('frac_of_host_peak') or 0, l['e2e']['d2h_bytes_per_step']/62.5e6))
print('per-rank launch ms', ['%.4f'%x for x in l['roofline']['launch_ms_mean_per_rank']])
if 'online_eval' in l:
  for sc,r in l['online_eval']['runs'].items():
    if isinstance(r,dict):
      for k,v in r.items(): print(sc, k, '%.1f k trajs/s frac %.3f envs/gpu %d'%(v['value']/1e3, v['roofline']['frac'], v['envs_per_gpu']))
"
if [ $N -gt 1 ]; then
timeout 600 python bench.py --no-cpu-baseline --no-other --no-online-eval --steps 50 > gpurun_out/bench_n1_same_box.json 2>/dev/null
python -c "
import json; l=json.load(open('gpurun_out/bench_n1_same_box.json')); print('N=1 same box: value %.1f G ms %.4f e2e %.2f G'%(l['value']/1e9, l['ms_per_step'], l['e2e']['value']/1e9))"
fi
