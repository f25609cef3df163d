#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_gp.py -q -m gpu -p no:cacheprovider -x 2>&1 | tail -15
timeout 600 python tools/bench_configs.py > gpurun_out/bench_configs_r33.jsonl 2>gpurun_out/bench_configs_r33.err; cut -c1-600 gpurun_out/bench_configs_r33.jsonl; tail -3 gpurun_out/bench_configs_r33.err
timeout 200 python tools/gp_once.py 8192 3 2>&1 | tail -1
