#!/bin/bash
# round 2, call 8: x3 with the cp.async weight ring: tests, cost by live batch, bench (default band and 3e-3), DenseNet flip study,
# ncu --set full of the two top kernels
mkdir -p gpurun_out
python -m pytest tests -m gpu -q -s -k "x3 or bench_config or vgp or engine" > gpurun_out/r02_pytest_8.log 2>&1; echo "pytest rc=$?"; grep -E "passed|failed|FAILED|Error|bench-config" gpurun_out/r02_pytest_8.log | cut -c1-300 | tail -20
timeout 300 python tools/r02_diag3.py resnet101 > gpurun_out/r02_diag3.log 2>&1; echo "diag3 rc=$?"; grep -A9 x3_ms gpurun_out/r02_diag3.log
timeout 600 python bench.py --no-cpu-baseline --no-gp --no-library-bar > gpurun_out/r02_bench_d.json 2> gpurun_out/r02_bench_d.err; echo "bench rc=$?"; cut -c1-300 gpurun_out/r02_bench_d.json; tail -3 gpurun_out/r02_bench_d.err
timeout 600 python bench.py --no-cpu-baseline --no-gp --no-library-bar --refine-ties 0.003 > gpurun_out/r02_bench_d_band3e-3.json 2>> gpurun_out/r02_bench_d.err; echo "bench3 rc=$?"; cut -c1-300 gpurun_out/r02_bench_d_band3e-3.json
timeout 600 python tools/r02_diag2.py 8192 densenet121 > gpurun_out/r02_diag2_densenet.log 2>&1; echo "diag2 densenet rc=$?"; head -50 gpurun_out/r02_diag2_densenet.log
NC="python bench.py --steps 1 --warmup 3 --masks-per-step 384 --micro-batch 384 --streams 1 --no-cpu-baseline --no-gp --no-library-bar --refine-ties 0"
$NC > gpurun_out/nc_plain.log 2>&1 && timeout 900 ncu --set full --clock-control none --import-source on -k regex:conv_fused_ca_kernel -s 60 -c 3 -o gpurun_out/r02_prof_fused $NC > gpurun_out/ncu_fused.log 2>&1; echo "ncu fused rc=$?"
$NC > gpurun_out/nc_plain2.log 2>&1 && timeout 900 ncu --set full --clock-control none --import-source on -k regex:conv_tc3_kernel -s 120 -c 4 -o gpurun_out/r02_prof_tc3 $NC > gpurun_out/ncu_tc3.log 2>&1; echo "ncu tc3 rc=$?"
ls -la gpurun_out/*.ncu-rep
