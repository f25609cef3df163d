#!/bin/bash
# round 2, call 16: 3x3 64->64 on the resident input patch (im2col mode 4): parity with the descriptor base offset on / off, role timers, bench
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_classifier.py -q -x -k "test_conv_tcgen05_vs_torch" > gpurun_out/r02_pytest_16_bo1.log 2>&1; echo "bo=1 pytest rc=$?"; tail -4 gpurun_out/r02_pytest_16_bo1.log
NIB_TC_HALO_BO=0 timeout 300 python -m pytest tests/test_gpu_classifier.py -q -k "test_conv_tcgen05_vs_torch" > gpurun_out/r02_pytest_16_bo0.log 2>&1; echo "bo=0 pytest rc=$?"; tail -4 gpurun_out/r02_pytest_16_bo0.log
NIB_TC_DBG=1 timeout 300 python bench.py --steps 1 --warmup 1 --masks-per-step 384 --micro-batch 384 --streams 1 --no-cpu-baseline --no-gp --no-library-bar --refine-ties 0 > /dev/null 2> gpurun_out/r02_role_timers_16.txt; grep "56x56 64->64 k3" gpurun_out/r02_role_timers_16.txt | tail -1
timeout 600 python bench.py --no-cpu-baseline --no-library-bar --no-gp --profile-json gpurun_out/r02_per_op_profile_16.json > gpurun_out/r02_bench_16.json 2> gpurun_out/r02_bench_16.err; echo "bench rc=$?"; cut -c1-300 gpurun_out/r02_bench_16.json
NIB_TC_NO_HALO=1 timeout 600 python bench.py --no-cpu-baseline --no-library-bar --no-gp > gpurun_out/r02_bench_16_nohalo.json 2>> gpurun_out/r02_bench_16.err; cut -c1-200 gpurun_out/r02_bench_16_nohalo.json
python tools/prof_table.py gpurun_out/r02_per_op_profile_16.json 2>/dev/null | head -12
