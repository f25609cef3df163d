#!/bin/bash
mkdir -p gpurun_out
nvidia-smi -L | head -3
timeout 600 python -m pytest tests/test_gpu_engine.py -q -m gpu -p no:cacheprovider -k two_rank 2>&1 | tail -3
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 5 --warmup 3 --no-gp > gpurun_out/bench_v4_n2.json 2>gpurun_out/bench_v4_n2.err; echo "rc=$?"; cut -c1-260 gpurun_out/bench_v4_n2.json; tail -2 gpurun_out/bench_v4_n2.err
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --impl reference --gpus 2 --steps 1 --warmup 1 2>/dev/null | cut -c1-200
