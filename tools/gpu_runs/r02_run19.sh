#!/bin/bash
# round 2, call 19: 3x3 128->128 on a resident patch (mode 6): conv parity, whole networks, role timers, A/B bench
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_classifier.py -q > gpurun_out/r02_pytest_19.log 2>&1; echo "classifier pytest rc=$?"; tail -6 gpurun_out/r02_pytest_19.log
NIB_TC_DBG=1 timeout 300 python bench.py --steps 1 --warmup 1 --masks-per-step 384 --micro-batch 384 --streams 1 --no-cpu-baseline --no-gp --no-library-bar --refine-ties 0 > /dev/null 2> gpurun_out/r02_role_timers_19.txt; grep "28x28 128->128 k3 s1" gpurun_out/r02_role_timers_19.txt | tail -1
timeout 600 python bench.py --no-cpu-baseline --no-library-bar --no-gp --profile-json gpurun_out/r02_per_op_profile_19.json > gpurun_out/r02_bench_19.json 2> gpurun_out/r02_bench_19.err; echo "bench rc=$?"; cut -c1-250 gpurun_out/r02_bench_19.json
NIB_TC_NO_HALO128=1 timeout 600 python bench.py --no-cpu-baseline --no-library-bar --no-gp > gpurun_out/r02_bench_19_no128.json 2>> gpurun_out/r02_bench_19.err; cut -c1-200 gpurun_out/r02_bench_19_no128.json
python tools/prof_table.py gpurun_out/r02_per_op_profile_19.json 2>/dev/null | head -9
