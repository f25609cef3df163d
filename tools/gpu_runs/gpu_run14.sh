#!/bin/bash
# 2-GPU: full GPU suite (incl. the torchrun sharding test), bench at N=2 and N=1, DenseNet-121 bench
mkdir -p gpurun_out
nvidia-smi -L
timeout 1500 python -m pytest tests -q -m gpu -x -p no:cacheprovider > gpurun_out/pytest_gpu_full2.log 2>&1
echo "pytest rc=$?"; tail -4 gpurun_out/pytest_gpu_full2.log
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 5 --warmup 3 > gpurun_out/bench_n2.log 2> gpurun_out/bench_n2.err
echo "bench n2 rc=$?"; tail -c 700 gpurun_out/bench_n2.log | head -c 400; echo
timeout 600 python bench.py --gpus 1 --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/bench_n1.log 2> gpurun_out/bench_n1.err
echo "bench n1 rc=$?"; head -c 300 gpurun_out/bench_n1.log; echo
timeout 600 python bench.py --arch densenet121 --steps 3 --warmup 3 --no-cpu-baseline --no-gp --profile-json gpurun_out/profile_densenet121.json > gpurun_out/bench_densenet121.log 2> gpurun_out/bench_densenet121.err
echo "densenet rc=$?"; head -c 300 gpurun_out/bench_densenet121.log; echo; tail -3 gpurun_out/bench_densenet121.err
