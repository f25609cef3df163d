#!/bin/bash
# round 2, call 17: stem on raw input rows (im2col mode 5, unswizzled overlapping-row descriptor) + 3x3 halo mode with base offset 0
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_classifier.py -q > gpurun_out/r02_pytest_17.log 2>&1; echo "classifier pytest rc=$?"; tail -6 gpurun_out/r02_pytest_17.log
NIB_TC_STEM_SWAP=1 timeout 600 python -m pytest tests/test_gpu_classifier.py -q -k "stem or whole or tv_resnet or resnet" > gpurun_out/r02_pytest_17_swap.log 2>&1; echo "swap pytest rc=$?"; tail -4 gpurun_out/r02_pytest_17_swap.log
NIB_TC_DBG=1 timeout 300 python bench.py --steps 1 --warmup 1 --masks-per-step 384 --micro-batch 384 --streams 1 --no-cpu-baseline --no-gp --no-library-bar --refine-ties 0 > /dev/null 2> gpurun_out/r02_role_timers_17.txt; grep "112x112" gpurun_out/r02_role_timers_17.txt | tail -1
timeout 600 python bench.py --no-cpu-baseline --no-library-bar --no-gp --profile-json gpurun_out/r02_per_op_profile_17.json > gpurun_out/r02_bench_17.json 2> gpurun_out/r02_bench_17.err; echo "bench rc=$?"; cut -c1-300 gpurun_out/r02_bench_17.json
NIB_TC_NO_HALO=1 timeout 600 python bench.py --no-cpu-baseline --no-library-bar --no-gp > gpurun_out/r02_bench_17_nohalo.json 2>> gpurun_out/r02_bench_17.err; cut -c1-200 gpurun_out/r02_bench_17_nohalo.json
python tools/prof_table.py gpurun_out/r02_per_op_profile_17.json 2>/dev/null | head -12
