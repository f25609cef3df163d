#!/bin/bash
# round 2, call 12: wavefront TRSV (A/B against the per-block launches), GP tests, full suite, bench; ncu source capture of potrf_diag
mkdir -p gpurun_out
python -m pytest tests/test_gpu_gp.py -q -x > gpurun_out/r02_pytest_gp_12.log 2>&1; echo "gp pytest rc=$?"; tail -3 gpurun_out/r02_pytest_gp_12.log
python tools/gp_once.py 8192 3 > gpurun_out/r02_gp_once_wave.txt 2>&1; cat gpurun_out/r02_gp_once_wave.txt
NIB_GP_TRSV_STEPS=1 python tools/gp_once.py 8192 3 > gpurun_out/r02_gp_once_steps.txt 2>&1; cat gpurun_out/r02_gp_once_steps.txt
python tools/gp_profile.py > gpurun_out/r02_gp_profile_12.txt 2>&1; tail -3 gpurun_out/r02_gp_profile_12.txt | cut -c1-600
python -m pytest tests -m gpu -q > gpurun_out/r02_pytest_12.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/r02_pytest_12.log
timeout 600 python bench.py > gpurun_out/r02_bench_12.json 2> gpurun_out/r02_bench_12.err; echo "bench rc=$?"; cut -c1-400 gpurun_out/r02_bench_12.json
python tools/r02_x3_probe.py 80 split > gpurun_out/r02_split_probe_12.txt 2>&1; tail -5 gpurun_out/r02_split_probe_12.txt
python tools/gp_once.py 2048 1 > /dev/null 2>&1 && timeout 600 ncu --set full --clock-control none --import-source on -k regex:potrf_diag_kernel -s 3 -c 1 -o gpurun_out/r02_potrf_diag python tools/gp_once.py 2048 1 > gpurun_out/ncu_potrf.log 2>&1; echo "ncu potrf rc=$?"
python tools/gp_once.py 8192 1 > /dev/null 2>&1 && timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 4000 --csv --log-file gpurun_out/r02_gp_launches_12.csv python tools/gp_once.py 8192 1 > gpurun_out/ncu_gp.log 2>&1; echo "ncu gp rc=$?"
python tools/agg_launches.py gpurun_out/r02_gp_launches_12.csv 14
