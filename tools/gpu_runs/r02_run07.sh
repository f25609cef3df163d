#!/bin/bash
# round 2, call 7: full pytest -m gpu with the split-bf16 re-score path; x3 / fp32 cost by live batch; default bench
mkdir -p gpurun_out
python -m pytest tests -m gpu -q -s > gpurun_out/r02_pytest_7.log 2>&1; echo "pytest rc=$?"; grep -E "passed|failed|FAILED|Error|bench-config|resnet56 bf16" gpurun_out/r02_pytest_7.log | cut -c1-300 | tail -30
timeout 300 python tools/r02_diag3.py resnet101 > gpurun_out/r02_diag3.log 2>&1; echo "diag3 rc=$?"; tail -40 gpurun_out/r02_diag3.log
timeout 600 python bench.py > gpurun_out/r02_bench_c.json 2> gpurun_out/r02_bench_c.err; echo "bench rc=$?"; cut -c1-700 gpurun_out/r02_bench_c.json; tail -3 gpurun_out/r02_bench_c.err
timeout 600 python bench.py --arch densenet121 --images 4 --masks-per-image 2048 --steps 2 --warmup 1 --no-cpu-baseline --no-gp --no-library-bar > gpurun_out/r02_bench_c_densenet_4img.json 2>> gpurun_out/r02_bench_c.err; echo "densenet rc=$?"; cut -c1-300 gpurun_out/r02_bench_c_densenet_4img.json
