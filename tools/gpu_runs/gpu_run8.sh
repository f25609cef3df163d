#!/bin/bash
# full GPU suite + smoke + default bench line (CPU baseline and GP included)
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.max.sm,clocks.sm --format=csv > gpurun_out/gpu.txt 2>&1
timeout 1500 python -m pytest tests -q -m gpu -x -p no:cacheprovider > gpurun_out/pytest_gpu_full.log 2>&1
echo "pytest rc=$?" >> gpurun_out/pytest_gpu_full.log
tail -5 gpurun_out/pytest_gpu_full.log
timeout 600 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?"; tail -2 gpurun_out/smoke.log
timeout 900 python bench.py --profile-json gpurun_out/profile_r8.json > gpurun_out/bench_r8.log 2> gpurun_out/bench_r8.err; echo "bench rc=$?"; tail -c 1500 gpurun_out/bench_r8.log
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_ref.log 2> gpurun_out/bench_ref.err; echo "ref rc=$?"; tail -c 600 gpurun_out/bench_ref.log
