#!/bin/bash
# round 2, 8-GPU call, shipped state: weak-scaling bench (driver's command), the literal configs[2] job (strong), configs[4] DenseNet, 2-GPU sharding test
mkdir -p gpurun_out
T="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29544"
timeout 600 $T bench.py --gpus 8 --steps 5 --warmup 3 --no-gp > gpurun_out/r02_bench_final_n8.json 2> gpurun_out/r02_bench_final_n8.err; echo "n8 rc=$?"; cut -c1-600 gpurun_out/r02_bench_final_n8.json; tail -3 gpurun_out/r02_bench_final_n8.err
timeout 600 $T bench.py --gpus 8 --strong --total-masks 16384 --steps 5 --warmup 3 --no-gp --no-cpu-baseline --no-library-bar > gpurun_out/r02_bench_final_n8_strong.json 2>> gpurun_out/r02_bench_final_n8.err; echo "n8 strong rc=$?"; cut -c1-300 gpurun_out/r02_bench_final_n8_strong.json
timeout 900 $T bench.py --gpus 8 --arch densenet121 --images 64 --masks-per-image 4096 --steps 3 --warmup 3 --no-gp --no-cpu-baseline --no-library-bar > gpurun_out/r02_bench_final_n8_densenet_config5.json 2>> gpurun_out/r02_bench_final_n8.err; echo "n8 densenet rc=$?"; cut -c1-300 gpurun_out/r02_bench_final_n8_densenet_config5.json
T4="python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29545"
timeout 600 $T4 bench.py --gpus 4 --steps 5 --warmup 3 --no-gp --no-cpu-baseline --no-library-bar > gpurun_out/r02_bench_final_n4.json 2>> gpurun_out/r02_bench_final_n8.err; echo "n4 rc=$?"; cut -c1-200 gpurun_out/r02_bench_final_n4.json
T2="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29546"
timeout 600 $T2 bench.py --gpus 2 --steps 5 --warmup 3 --no-gp --no-cpu-baseline --no-library-bar > gpurun_out/r02_bench_final_n2.json 2>> gpurun_out/r02_bench_final_n8.err; echo "n2 rc=$?"; cut -c1-200 gpurun_out/r02_bench_final_n2.json
timeout 600 python bench.py --steps 5 --warmup 3 --no-gp --no-cpu-baseline --no-library-bar > gpurun_out/r02_bench_final_n1_same_box.json 2>> gpurun_out/r02_bench_final_n8.err; echo "n1 rc=$?"; cut -c1-200 gpurun_out/r02_bench_final_n1_same_box.json
timeout 600 python -m pytest tests -m gpu -q -k "multi_gpu or sharded or two_gpu or 2gpu" > gpurun_out/r02_pytest_final_2gpu.log 2>&1; echo "2gpu pytest rc=$?"; tail -3 gpurun_out/r02_pytest_final_2gpu.log
