mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
tail -15 gpurun_out/pytest_gpu.log
python scripts/bench_kernels.py --reps 20 > gpurun_out/kernels.jsonl 2> gpurun_out/kernels.err; echo "kernels rc=$?"
tail -3 gpurun_out/kernels.err
cut -c1-330 gpurun_out/kernels.jsonl
