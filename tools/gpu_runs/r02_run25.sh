#!/bin/bash
# round 2, call 25: DenseNet-121 per-op profile + launch list (where does its time go)
mkdir -p gpurun_out
timeout 600 python bench.py --arch densenet121 --images 2 --masks-per-image 3072 --no-cpu-baseline --no-library-bar --no-gp --refine-ties 0 --profile-json gpurun_out/r02_per_op_profile_densenet121.json > gpurun_out/r02_bench_25_densenet.json 2> gpurun_out/r02_bench_25.err; echo "bench rc=$?"; cut -c1-250 gpurun_out/r02_bench_25_densenet.json
python tools/prof_table.py gpurun_out/r02_per_op_profile_densenet121.json 2>/dev/null | head -40
D="python bench.py --arch densenet121 --images 1 --masks-per-image 1536 --steps 1 --warmup 3 --no-cpu-baseline --no-gp --no-library-bar --refine-ties 0"
$D > /dev/null 2>&1 && timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file gpurun_out/r02_launches_densenet.csv $D > gpurun_out/ncu_densenet.log 2>&1; echo "ncu rc=$?"
python tools/agg_launches.py gpurun_out/r02_launches_densenet.csv 14
