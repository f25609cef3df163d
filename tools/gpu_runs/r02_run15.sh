#!/bin/bash
# round 2, call 15: tensor-core fc + branch-free max pool: classifier/engine tests, bench, per-op profile; role timers of every tcgen05 launch
mkdir -p gpurun_out
python -m pytest tests/test_gpu_classifier.py tests/test_gpu_engine.py tests/test_gpu_bench_config.py -q -x > gpurun_out/r02_pytest_15.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/r02_pytest_15.log
timeout 600 python bench.py --no-cpu-baseline --no-library-bar --profile-json gpurun_out/r02_per_op_profile_15.json > gpurun_out/r02_bench_15.json 2> gpurun_out/r02_bench_15.err; echo "bench rc=$?"; cut -c1-300 gpurun_out/r02_bench_15.json
NIB_FC_SIMT=1 timeout 600 python bench.py --no-cpu-baseline --no-library-bar --no-gp > gpurun_out/r02_bench_15_fcsimt.json 2>> gpurun_out/r02_bench_15.err; cut -c1-200 gpurun_out/r02_bench_15_fcsimt.json
python tools/prof_table.py gpurun_out/r02_per_op_profile_15.json | head -14
NIB_TC_DBG=1 python bench.py --steps 1 --warmup 1 --masks-per-step 384 --micro-batch 384 --streams 1 --no-cpu-baseline --no-gp --no-library-bar --refine-ties 0 > /dev/null 2> gpurun_out/r02_role_timers_15.txt; grep -c tc3 gpurun_out/r02_role_timers_15.txt
