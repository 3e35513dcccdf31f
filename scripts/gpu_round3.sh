mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
tail -40 gpurun_out/pytest_gpu.log
python scripts/bench_kernels.py --reps 5 --only gpt2 > gpurun_out/kernels_gpt2.jsonl 2> gpurun_out/kernels_gpt2.err; echo "kernels rc=$?"
tail -5 gpurun_out/kernels_gpt2.err
cut -c1-400 gpurun_out/kernels_gpt2.jsonl
