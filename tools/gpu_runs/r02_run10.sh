#!/bin/bash
# round 2, call 10: split mode of the pair kernel: conv-level tests, whole nets, tie policy, then the rest of the suite and the bench
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_classifier.py -m gpu -q -x -k "split" > gpurun_out/r02_pytest_10a.log 2>&1; echo "split tests rc=$?"; tail -25 gpurun_out/r02_pytest_10a.log | cut -c1-250
timeout 900 python -m pytest tests -m gpu -q > gpurun_out/r02_pytest_10.log 2>&1; echo "pytest rc=$?"; grep -E "passed|failed|FAILED|Error" gpurun_out/r02_pytest_10.log | cut -c1-250 | tail -15
timeout 300 python tools/r02_x3_probe.py 80 split > gpurun_out/r02_split_probe_80.log 2>&1; echo "probe rc=$?"; tail -16 gpurun_out/r02_split_probe_80.log
timeout 300 python tools/r02_x3_probe.py 256 split > gpurun_out/r02_split_probe_256.log 2>&1; head -3 gpurun_out/r02_split_probe_256.log | tail -2
timeout 600 python bench.py --no-cpu-baseline --no-gp --no-library-bar > gpurun_out/r02_bench_e.json 2> gpurun_out/r02_bench_e.err; echo "bench rc=$?"; cut -c1-300 gpurun_out/r02_bench_e.json; tail -3 gpurun_out/r02_bench_e.err
