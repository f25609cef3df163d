#!/bin/bash
# round 2, call 27: rows-per-tile = largest divisor of the height (14x14 growth convs join the resident-patch mode): parity, DenseNet bench + profile
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_classifier.py tests/test_gpu_engine.py -q > gpurun_out/r02_pytest_27.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/r02_pytest_27.log
timeout 600 python bench.py --arch densenet121 --images 2 --masks-per-image 3072 --no-cpu-baseline --no-library-bar --no-gp --refine-ties 0 --profile-json gpurun_out/r02_per_op_profile_densenet121.json > gpurun_out/r02_bench_27_densenet.json 2> gpurun_out/r02_bench_27.err; echo "bench rc=$?"; cut -c1-200 gpurun_out/r02_bench_27_densenet.json
python tools/prof_table.py gpurun_out/r02_per_op_profile_densenet121.json 2>/dev/null | head -6
timeout 600 python bench.py --arch densenet121 --images 8 --masks-per-image 4096 --no-cpu-baseline --no-library-bar --no-gp > gpurun_out/r02_bench_27_densenet_tie.json 2>> gpurun_out/r02_bench_27.err; python -c "
import json; d=json.loads(open('gpurun_out/r02_bench_27_densenet_tie.json').read().strip().splitlines()[-1]); print('densenet tie on', d['value'], d.get('near_ties_per_step'), d.get('tie_overflow_per_step'))"
D="python bench.py --arch densenet121 --images 1 --masks-per-image 1536 --steps 1 --warmup 3 --no-cpu-baseline --no-gp --no-library-bar --refine-ties 0"
$D > /dev/null 2>&1 && timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file gpurun_out/r02_launches_densenet.csv $D > gpurun_out/ncu_densenet.log 2>&1; echo "ncu rc=$?"
python tools/agg_launches.py gpurun_out/r02_launches_densenet.csv 8
