#!/bin/bash
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
tail -4 gpurun_out/pytest_gpu.log
for cfg in "1 10000 tma_c4" "1 1250 tma_c4_8gpu" "0 10000 regstaged_c4"; do set -- $cfg
  DPT_GPT2_TMA=$1 DPT_PROFILE_ENVS=$2 timeout 400 ncu --set full --clock-control none --import-source on -k regex:gpt2_online --launch-skip 1 --launch-count 1 -f -o gpurun_out/r02_gpt2_fp32_$3 python scripts/profile_target.py gpt2_c4 2 > gpurun_out/ncu_$3.log 2>&1; echo "ncu $3 rc=$?"
done
