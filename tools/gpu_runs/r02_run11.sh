#!/bin/bash
# round 2, call 11: smoke(), full pytest -m gpu, the driver's two bench arms, launch list of the default command (shipped state)
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r02_smoke.log 2>&1; echo "smoke rc=$?"; tail -2 gpurun_out/r02_smoke.log | cut -c1-400
python -m pytest tests -m gpu -q > gpurun_out/r02_pytest_11.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/r02_pytest_11.log
timeout 300 python bench.py --impl reference --gpus 1 --steps 3 --warmup 1 > gpurun_out/r02_bench_f_ref.json 2> gpurun_out/r02_bench_f.err; cut -c1-200 gpurun_out/r02_bench_f_ref.json
timeout 600 python bench.py --profile-json gpurun_out/r02_per_op_profile_mb384.json > gpurun_out/r02_bench_f.json 2>> gpurun_out/r02_bench_f.err; echo "bench rc=$?"; cut -c1-300 gpurun_out/r02_bench_f.json; tail -3 gpurun_out/r02_bench_f.err
D="python bench.py --no-cpu-baseline --no-gp --no-library-bar"
$D > gpurun_out/d_plain.log 2>&1 && timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 3200 --csv --log-file gpurun_out/r02_launches_final.csv $D > gpurun_out/ncu_default.log 2>&1; echo "ncu default rc=$? lines=$(wc -l < gpurun_out/r02_launches_final.csv)"
python tools/agg_launches.py gpurun_out/r02_launches_final.csv 18
