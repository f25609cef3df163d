#!/bin/bash
# round 2, call 22: fused kernel with two staging sets per chunk parity (N2 <= 128), tie capacity 1024: parity, stress, per-op profile, bench, DenseNet overflow
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_classifier.py tests/test_gpu_engine.py tests/test_gpu_bench_config.py -q -x > gpurun_out/r02_pytest_22.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/r02_pytest_22.log
timeout 600 python tools/fused_stress.py > gpurun_out/r02_fused_stress_22.txt 2>&1; echo "stress rc=$?"; tail -4 gpurun_out/r02_fused_stress_22.txt | cut -c1-300
NIB_TC_DBG=1 timeout 300 python bench.py --steps 1 --warmup 1 --masks-per-step 384 --micro-batch 384 --streams 1 --no-cpu-baseline --no-gp --no-library-bar --refine-ties 0 > /dev/null 2> gpurun_out/r02_role_timers_22.txt; grep -A8 "fused K1=64 " gpurun_out/r02_role_timers_22.txt | head -9
timeout 600 python bench.py --no-cpu-baseline --no-library-bar --no-gp --profile-json gpurun_out/r02_per_op_profile_22.json > gpurun_out/r02_bench_22.json 2> gpurun_out/r02_bench_22.err; echo "bench rc=$?"; cut -c1-250 gpurun_out/r02_bench_22.json
python tools/prof_table.py gpurun_out/r02_per_op_profile_22.json 2>/dev/null | head -9
timeout 600 python bench.py --arch densenet121 --images 8 --masks-per-image 4096 --no-cpu-baseline --no-library-bar --no-gp > gpurun_out/r02_bench_22_densenet.json 2>> gpurun_out/r02_bench_22.err; python -c "
import json; d=json.loads(open('gpurun_out/r02_bench_22_densenet.json').read().strip().splitlines()[-1]); print('densenet', d['value'], d.get('near_ties_per_step'), d.get('tie_overflow_per_step'))"
nvidia-smi --query-gpu=memory.used --format=csv | tail -1
