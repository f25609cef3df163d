#!/bin/bash
# round 2, call 18: shipped state after the halo modes: smoke, full pytest -m gpu, both bench arms, launch list of the default command, DenseNet
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r02_smoke.log 2>&1; echo "smoke rc=$?"; tail -2 gpurun_out/r02_smoke.log | cut -c1-400
python -m pytest tests -m gpu -q > gpurun_out/r02_pytest_18.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/r02_pytest_18.log
timeout 300 python bench.py --impl reference --gpus 1 --steps 3 --warmup 1 > gpurun_out/r02_bench_18_ref.json 2> gpurun_out/r02_bench_18.err; cut -c1-200 gpurun_out/r02_bench_18_ref.json
timeout 600 python bench.py --profile-json gpurun_out/r02_per_op_profile_18.json > gpurun_out/r02_bench_18.json 2>> gpurun_out/r02_bench_18.err; echo "bench rc=$?"; cut -c1-300 gpurun_out/r02_bench_18.json; tail -3 gpurun_out/r02_bench_18.err
timeout 600 python bench.py --no-cpu-baseline --no-library-bar --no-gp --refine-ties 0 > gpurun_out/r02_bench_18_tieoff.json 2>> gpurun_out/r02_bench_18.err; cut -c1-200 gpurun_out/r02_bench_18_tieoff.json
timeout 600 python bench.py --arch densenet121 --images 8 --masks-per-image 4096 --no-cpu-baseline --no-library-bar --no-gp > gpurun_out/r02_bench_18_densenet.json 2>> gpurun_out/r02_bench_18.err; cut -c1-200 gpurun_out/r02_bench_18_densenet.json
D="python bench.py --no-cpu-baseline --no-gp --no-library-bar"
$D > gpurun_out/d_plain.log 2>&1 && timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 3200 --csv --log-file gpurun_out/r02_launches_18.csv $D > gpurun_out/ncu_default.log 2>&1; echo "ncu default rc=$? lines=$(wc -l < gpurun_out/r02_launches_18.csv)"
python tools/agg_launches.py gpurun_out/r02_launches_18.csv 18
