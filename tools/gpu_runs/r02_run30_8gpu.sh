#!/bin/bash
# round 2, final 8-GPU call: the driver's weak-scaling command, the literal configs[2] job, configs[4] DenseNet (tie policy on and off), 1 GPU of the same box
mkdir -p gpurun_out
T="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29544"
timeout 600 $T bench.py --gpus 8 --steps 5 --warmup 3 --no-gp > gpurun_out/r02_bench_final_n8.json 2> gpurun_out/r02_bench_final_n8.err; echo "n8 rc=$?"; grep '^{' gpurun_out/r02_bench_final_n8.json | cut -c1-200
timeout 600 $T bench.py --gpus 8 --strong --total-masks 16384 --steps 5 --warmup 3 --no-gp --no-cpu-baseline --no-library-bar > gpurun_out/r02_bench_final_n8_strong.json 2>> gpurun_out/r02_bench_final_n8.err; echo "n8 strong rc=$?"; grep '^{' gpurun_out/r02_bench_final_n8_strong.json | cut -c1-200
timeout 900 $T bench.py --gpus 8 --arch densenet121 --images 64 --masks-per-image 4096 --steps 3 --warmup 3 --no-gp --no-cpu-baseline --no-library-bar > gpurun_out/r02_bench_final_n8_densenet_config5.json 2>> gpurun_out/r02_bench_final_n8.err; echo "n8 densenet rc=$?"; grep '^{' gpurun_out/r02_bench_final_n8_densenet_config5.json | cut -c1-200
timeout 900 $T bench.py --gpus 8 --arch densenet121 --images 64 --masks-per-image 4096 --steps 3 --warmup 3 --no-gp --no-cpu-baseline --no-library-bar --refine-ties 0 > gpurun_out/r02_bench_final_n8_densenet_config5_tie_off.json 2>> gpurun_out/r02_bench_final_n8.err; echo "n8 densenet tie off rc=$?"; grep '^{' gpurun_out/r02_bench_final_n8_densenet_config5_tie_off.json | cut -c1-200
timeout 600 python bench.py --steps 5 --warmup 3 --no-gp --no-cpu-baseline --no-library-bar > gpurun_out/r02_bench_final_n1_same_box.json 2>> gpurun_out/r02_bench_final_n8.err; echo "n1 rc=$?"; cut -c1-200 gpurun_out/r02_bench_final_n1_same_box.json
