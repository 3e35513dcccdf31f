#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_online_gpu.py -m gpu -x -q > gpurun_out/pytest_online.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_online.log
tail -8 gpurun_out/pytest_online.log
V=decision-pretrained-transformer_b200/variants
for v in default wt16; do
  if [ $v = default ]; then unset DPT_B200_LIB; else export DPT_B200_LIB=$PWD/$V/libdpt_b200_$v.so; fi
  timeout 300 python scripts/bench_kernels.py --only online > gpurun_out/k_online_$v.jsonl 2> gpurun_out/k_online_$v.err; echo "$v rc=$?"
done
unset DPT_B200_LIB
python - <<'PY'
import json
for f in ("default","wt16"):
    for l in open("gpurun_out/k_online_%s.jsonl"%f):
        j=json.loads(l); print(f, j["kernel"], "%.3f ms"%j["ms_mean"], "frac %.3f"%j["frac_of_measured_hbm_peak"])
PY
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_online_emp.csv python scripts/profile_target.py online_emp 3 > gpurun_out/ncu_l.log 2>&1; echo "rc=$?"
grep -E "online_loop|regret|table|reduce|fill" gpurun_out/launches_online_emp.csv | awk -F'","' '{print $5, $NF}' | tail -4
