#!/bin/bash
# round 2, call 24: fused kernel with per-K-block hand-over of the resident A tile: parity, stress, role timers, A/B bench on one box (git stash of the .so is not possible: A/B by timers)
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_classifier.py tests/test_gpu_engine.py -q -x > gpurun_out/r02_pytest_24.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/r02_pytest_24.log
timeout 600 python tools/fused_stress.py > gpurun_out/r02_fused_stress_24.txt 2>&1; echo "stress rc=$?"; tail -2 gpurun_out/r02_fused_stress_24.txt | cut -c1-300
NIB_TC_DBG=1 timeout 300 python bench.py --steps 1 --warmup 1 --masks-per-step 384 --micro-batch 384 --streams 1 --no-cpu-baseline --no-gp --no-library-bar --refine-ties 0 > /dev/null 2> gpurun_out/r02_role_timers_24.txt; grep -A8 "fused K1=256 " gpurun_out/r02_role_timers_24.txt | tail -9; grep "14x14 256->256 k3 s1" gpurun_out/r02_role_timers_24.txt | tail -1 | cut -c1-120
for i in 1 2; do timeout 600 python bench.py --no-cpu-baseline --no-library-bar --no-gp --profile-json gpurun_out/r02_per_op_profile_24.json > gpurun_out/r02_bench_24.json 2> gpurun_out/r02_bench_24.err; cut -c1-200 gpurun_out/r02_bench_24.json; done
python tools/prof_table.py gpurun_out/r02_per_op_profile_24.json 2>/dev/null | head -6
