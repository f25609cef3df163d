#!/bin/bash
mkdir -p gpurun_out
n=8
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $((29530+n)) bench.py --gpus $n --steps 5 --warmup 3 --no-gp > gpurun_out/bench_v6_n$n.json 2>gpurun_out/bench_v6_n$n.err; echo "n=$n rc=$?"; grep '"metric"' gpurun_out/bench_v6_n$n.json | cut -c1-230
