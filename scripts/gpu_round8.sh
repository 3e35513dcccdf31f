mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
tail -5 gpurun_out/pytest_gpu.log
for t in gpt2 gpt2_bf16; do
  python scripts/profile_target.py $t 3 > gpurun_out/plain_$t.log 2>&1 && \
  ncu --set full --clock-control none --import-source on -k regex:"gpt2_online" -s 1 -c 1 -o gpurun_out/prof_$t -f python scripts/profile_target.py $t 3 > gpurun_out/ncu_$t.log 2>&1
  echo "$t rc=$?"
done
