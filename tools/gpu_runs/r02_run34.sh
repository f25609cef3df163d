#!/bin/bash
# round 2, call 34: configs[0/1/3] with the full-window warm-up
mkdir -p gpurun_out
timeout 600 python tools/bench_configs.py > gpurun_out/r02_bench_configs_0_1_3.jsonl 2> gpurun_out/r02_bench_configs.err; echo "configs rc=$?"; grep "configs\[3\]" gpurun_out/r02_bench_configs_0_1_3.jsonl | cut -c1-400; tail -2 gpurun_out/r02_bench_configs.err
