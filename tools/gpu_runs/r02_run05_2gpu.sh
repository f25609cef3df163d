#!/bin/bash
# round 2, 2-GPU call: the torchrun sharded == unsharded test through the C-ABI all-gather, bench at N=2 (weak, strong)
mkdir -p gpurun_out
python -m pytest tests/test_gpu_engine.py -m gpu -q -k "two_rank" > gpurun_out/r02_pytest_2gpu.log 2>&1; echo "pytest 2gpu rc=$?"; tail -3 gpurun_out/r02_pytest_2gpu.log
T="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533"
timeout 600 $T bench.py --gpus 2 --no-gp > gpurun_out/r02_bench_n2.json 2> gpurun_out/r02_bench_n2.err; echo "n2 rc=$?"; cut -c1-1200 gpurun_out/r02_bench_n2.json; tail -3 gpurun_out/r02_bench_n2.err
timeout 600 $T bench.py --gpus 2 --strong --total-masks 16384 --steps 3 --warmup 1 --no-gp > gpurun_out/r02_bench_n2_strong.json 2>> gpurun_out/r02_bench_n2.err; echo "n2 strong rc=$?"; cut -c1-500 gpurun_out/r02_bench_n2_strong.json
