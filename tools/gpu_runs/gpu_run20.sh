#!/bin/bash
mkdir -p gpurun_out
timeout 1200 python tools/bench_configs.py > gpurun_out/bench_configs.jsonl 2> gpurun_out/bench_configs.err; echo "configs rc=$?"; cat gpurun_out/bench_configs.jsonl | cut -c1-400; tail -3 gpurun_out/bench_configs.err
