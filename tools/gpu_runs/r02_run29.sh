#!/bin/bash
# round 2, call 29: shipped state: smoke, full pytest -m gpu, both bench arms, DenseNet with / without the tie policy, launch lists
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r02_smoke.log 2>&1; echo "smoke rc=$?"; tail -2 gpurun_out/r02_smoke.log | cut -c1-400
python -m pytest tests -m gpu -q > gpurun_out/r02_pytest_29.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/r02_pytest_29.log
timeout 300 python bench.py --impl reference --gpus 1 --steps 3 --warmup 1 > gpurun_out/r02_bench_29_ref.json 2> gpurun_out/r02_bench_29.err; cut -c1-200 gpurun_out/r02_bench_29_ref.json
timeout 600 python bench.py --profile-json gpurun_out/r02_per_op_profile_29.json > gpurun_out/r02_bench_29.json 2>> gpurun_out/r02_bench_29.err; echo "bench rc=$?"; cut -c1-300 gpurun_out/r02_bench_29.json; tail -3 gpurun_out/r02_bench_29.err
timeout 600 python bench.py --no-cpu-baseline --no-library-bar --no-gp --refine-ties 0 > gpurun_out/r02_bench_29_tieoff.json 2>> gpurun_out/r02_bench_29.err; cut -c1-200 gpurun_out/r02_bench_29_tieoff.json
timeout 600 python bench.py --arch densenet121 --images 8 --masks-per-image 4096 --no-cpu-baseline --no-library-bar --no-gp > gpurun_out/r02_bench_29_densenet.json 2>> gpurun_out/r02_bench_29.err; python -c "
import json; d=json.loads(open('gpurun_out/r02_bench_29_densenet.json').read().strip().splitlines()[-1]); print('densenet tie on', d['value'], d.get('near_ties_per_step'), d.get('tie_overflow_per_step'))"
timeout 600 python bench.py --arch densenet121 --images 8 --masks-per-image 4096 --no-cpu-baseline --no-library-bar --no-gp --refine-ties 0 --profile-json gpurun_out/r02_per_op_profile_densenet121.json > gpurun_out/r02_bench_29_densenet_tieoff.json 2>> gpurun_out/r02_bench_29.err; cut -c1-200 gpurun_out/r02_bench_29_densenet_tieoff.json
D="python bench.py --no-cpu-baseline --no-gp --no-library-bar"
$D > gpurun_out/d_plain.log 2>&1 && timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 3200 --csv --log-file gpurun_out/r02_launches_29.csv $D > gpurun_out/ncu_default.log 2>&1; echo "ncu default rc=$? lines=$(wc -l < gpurun_out/r02_launches_29.csv)"
python tools/agg_launches.py gpurun_out/r02_launches_29.csv 8
DD="python bench.py --arch densenet121 --images 1 --masks-per-image 1536 --steps 1 --warmup 3 --no-cpu-baseline --no-gp --no-library-bar --refine-ties 0"
$DD > /dev/null 2>&1 && timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file gpurun_out/r02_launches_densenet_29.csv $DD > gpurun_out/ncu_densenet.log 2>&1; echo "ncu densenet rc=$?"
python tools/agg_launches.py gpurun_out/r02_launches_densenet_29.csv 8
