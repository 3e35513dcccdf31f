mkdir -p gpurun_out
for t in darkroom online_emp online_thompson gpt2; do
  python scripts/profile_target.py $t 3 > gpurun_out/plain_$t.log 2>&1 && \
  ncu --set full --clock-control none --import-source on -k regex:"rollin|online|gpt2" -s 1 -c 1 -o gpurun_out/prof_$t -f python scripts/profile_target.py $t 3 > gpurun_out/ncu_$t.log 2>&1
  echo "$t rc=$?"
done
python bench.py --gpus 1 --steps 200 --warmup 10 > gpurun_out/bench_n1.log 2>gpurun_out/bench_n1.err; echo "bench1 rc=$?"
cut -c1-400 gpurun_out/bench_n1.log
