#!/bin/bash
# round 2, call 4: full pytest -m gpu; default bench (tie band 4.5e-3); strong / multi-image modes; ncu launch lists + all-launch capture
mkdir -p gpurun_out
python -m pytest tests -m gpu -q -s > gpurun_out/r02_pytest_4.log 2>&1; echo "pytest rc=$?"; grep -E "passed|failed|FAILED|Error|bench-config|resnet56 bf16" gpurun_out/r02_pytest_4.log | cut -c1-300 | tail -30
timeout 600 python bench.py --profile-json gpurun_out/r02_per_op_profile_mb384.json > gpurun_out/r02_bench_b.json 2> gpurun_out/r02_bench_b.err; echo "bench rc=$?"; cut -c1-900 gpurun_out/r02_bench_b.json; tail -3 gpurun_out/r02_bench_b.err
timeout 300 python bench.py --strong --total-masks 16384 --steps 2 --warmup 1 --no-cpu-baseline --no-gp --no-library-bar > gpurun_out/r02_bench_b_strong.json 2>> gpurun_out/r02_bench_b.err; echo "strong rc=$?"; cut -c1-400 gpurun_out/r02_bench_b_strong.json
timeout 600 python bench.py --arch densenet121 --images 4 --masks-per-image 2048 --steps 2 --warmup 1 --no-cpu-baseline --no-gp --no-library-bar > gpurun_out/r02_bench_b_densenet_4img.json 2>> gpurun_out/r02_bench_b.err; echo "densenet rc=$?"; cut -c1-700 gpurun_out/r02_bench_b_densenet_4img.json
NC="python bench.py --steps 1 --warmup 3 --masks-per-step 384 --micro-batch 384 --streams 1 --no-cpu-baseline --no-gp --no-library-bar --refine-ties 0"
M="gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed,sm__cycles_active.avg,sm__cycles_elapsed.max,lts__t_sector_hit_rate.pct,l1tex__m_xbar2l1tex_read_bytes.sum"
$NC > gpurun_out/nc_plain.log 2>&1 && timeout 900 ncu --metrics $M --clock-control none --csv --page raw --log-file gpurun_out/r02_all_launches_raw.csv -s 250 -c 200 $NC > gpurun_out/ncu_all.log 2>&1; echo "ncu all rc=$? lines=$(wc -l < gpurun_out/r02_all_launches_raw.csv)"
python tools/ncu_all_launches.py gpurun_out/r02_all_launches_raw.csv gpurun_out/r02_ncu_all_launches_one_forward.csv
D="python bench.py --no-cpu-baseline --no-gp --no-library-bar"
$D > gpurun_out/d_plain.log 2>&1 && timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file gpurun_out/r02_launches_default.csv $D > gpurun_out/ncu_default.log 2>&1; echo "ncu default rc=$? lines=$(wc -l < gpurun_out/r02_launches_default.csv)"
python tools/agg_launches.py gpurun_out/r02_launches_default.csv 14
