#!/bin/bash
# round 2, call 31: BASELINE configs[0], [1] (fp32 / x3 / bf16), [3] through the package API
mkdir -p gpurun_out
timeout 600 python tools/bench_configs.py > gpurun_out/r02_bench_configs_0_1_3.jsonl 2> gpurun_out/r02_bench_configs.err; echo "rc=$?"; cut -c1-330 gpurun_out/r02_bench_configs_0_1_3.jsonl; tail -3 gpurun_out/r02_bench_configs.err
