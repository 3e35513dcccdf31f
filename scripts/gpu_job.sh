#!/bin/bash
# GPU job 1 (round 2): tests, bench N=1, ncu of the decode kernels at config 4
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.max.sm --format=csv > gpurun_out/gpu.txt
nproc > gpurun_out/nproc.txt
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
tail -3 gpurun_out/pytest_gpu.log
timeout 600 python bench.py > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err; echo "bench rc=$?"
tail -c 3000 gpurun_out/bench_n1.json
timeout 300 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err; echo "ref rc=$?"
for t in gpt2_c4 gpt2_bf16_c4; do
  timeout 400 ncu --set full --clock-control none --import-source on -k regex:gpt2_online --launch-skip 1 --launch-count 1 -f -o gpurun_out/r02_${t} python scripts/profile_target.py $t 2 > gpurun_out/ncu_${t}.log 2>&1; echo "ncu $t rc=$?"
done
ls -la gpurun_out | tail -20
