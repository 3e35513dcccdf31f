mkdir -p gpurun_out
for t in bandit darkroom dense_bf16 dense_fp32 online_thompson online_emp; do
  python scripts/profile_target.py $t 3 > gpurun_out/plain_$t.log 2>&1 && \
  ncu --set full --clock-control none --import-source on -k regex:"rollin|online|gpt2" -s 1 -c 1 -o gpurun_out/prof2_$t -f python scripts/profile_target.py $t 3 > gpurun_out/ncu2_$t.log 2>&1
  echo "$t rc=$?"
done
