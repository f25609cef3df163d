#!/bin/bash
# first GPU contact: per-geometry tcgen05 diagnostics, the full -m gpu suite, one short bench
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,memory.total --format=csv > gpurun_out/gpu.txt 2>&1
timeout 900 python tools/tc_diag.py > gpurun_out/tc_diag.log 2>&1
timeout 1500 python -m pytest tests -q -m gpu -p no:cacheprovider --timeout=600 > gpurun_out/pytest_gpu.log 2>&1
echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
timeout 600 python bench.py --steps 3 --warmup 3 > gpurun_out/bench.log 2> gpurun_out/bench.err
echo "bench rc=$?" >> gpurun_out/bench.err
tail -5 gpurun_out/pytest_gpu.log; cat gpurun_out/tc_diag.txt; tail -3 gpurun_out/bench.log
