#!/bin/bash
mkdir -p gpurun_out
NC="python bench.py --steps 1 --warmup 3 --masks-per-step 256 --micro-batch 256 --no-cpu-baseline --no-gp"
$NC > gpurun_out/plain.log 2>&1 && \
timeout 1500 ncu --set full --clock-control none --import-source on -k regex:conv_tc2_kernel -s 145 -c 7 -o gpurun_out/prof_conv_tc2 $NC > gpurun_out/ncu_full.log 2>&1
tail -3 gpurun_out/ncu_full.log | cut -c1-300
ls -la gpurun_out/prof_conv_tc2.ncu-rep
