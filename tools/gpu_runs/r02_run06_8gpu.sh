#!/bin/bash
# round 2, 8-GPU call: weak-scaling bench, the literal configs[2] job (16384 masks, strong), configs[4] (DenseNet-121, 64 x 4096)
mkdir -p gpurun_out
T="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29544"
timeout 600 $T bench.py --gpus 8 --no-gp > gpurun_out/r02_bench_n8.json 2> gpurun_out/r02_bench_n8.err; echo "n8 rc=$?"; cut -c1-1200 gpurun_out/r02_bench_n8.json; tail -3 gpurun_out/r02_bench_n8.err
timeout 600 $T bench.py --gpus 8 --strong --total-masks 16384 --steps 5 --warmup 3 --no-gp > gpurun_out/r02_bench_n8_strong.json 2>> gpurun_out/r02_bench_n8.err; echo "n8 strong rc=$?"; cut -c1-500 gpurun_out/r02_bench_n8_strong.json
timeout 900 $T bench.py --gpus 8 --arch densenet121 --images 64 --masks-per-image 4096 --steps 3 --warmup 1 --no-gp > gpurun_out/r02_bench_n8_densenet_config5.json 2>> gpurun_out/r02_bench_n8.err; echo "n8 densenet rc=$?"; cut -c1-700 gpurun_out/r02_bench_n8_densenet_config5.json
