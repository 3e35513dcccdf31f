mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
tail -5 gpurun_out/pytest_gpu.log
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 scripts/multi_gpu_check.py > gpurun_out/multi_gpu_check.log 2>&1; echo "mgpu rc=$?"
tail -3 gpurun_out/multi_gpu_check.log
python bench.py --gpus 1 --steps 100 --warmup 5 > gpurun_out/bench_n1.log 2>gpurun_out/bench_n1.err; echo "bench1 rc=$?"
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --steps 100 --warmup 5 > gpurun_out/bench_n2.log 2>gpurun_out/bench_n2.err; echo "bench2 rc=$?"
tail -2 gpurun_out/bench_n2.err
cut -c1-600 gpurun_out/bench_n1.log; cut -c1-900 gpurun_out/bench_n2.log
