#!/bin/bash
# round 2, call 13: lean potrf_diag pivot loop + sentinel-polling TRSV: GP tests, timings, launch list, ncu of potrf_diag
mkdir -p gpurun_out
python -m pytest tests/test_gpu_gp.py tests/test_gpu_vgp.py -q -x > gpurun_out/r02_pytest_gp_13.log 2>&1; echo "gp pytest rc=$?"; tail -3 gpurun_out/r02_pytest_gp_13.log
python tools/gp_once.py 8192 3 > gpurun_out/r02_gp_once_13.txt 2>&1; cat gpurun_out/r02_gp_once_13.txt
python tools/gp_once.py 1000 2; python tools/gp_once.py 4096 2
python tools/gp_profile.py > gpurun_out/r02_gp_profile_13.txt 2>&1; tail -3 gpurun_out/r02_gp_profile_13.txt | cut -c1-600
python tools/gp_once.py 8192 1 > /dev/null 2>&1 && timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 4000 --csv --log-file gpurun_out/r02_gp_launches_13.csv python tools/gp_once.py 8192 1 > gpurun_out/ncu_gp.log 2>&1; echo "ncu gp rc=$?"
python tools/agg_launches.py gpurun_out/r02_gp_launches_13.csv 8
timeout 600 ncu --set full --clock-control none --import-source on -k regex:potrf_diag_kernel -s 3 -c 1 -o gpurun_out/r02_potrf_diag_13 python tools/gp_once.py 2048 1 > gpurun_out/ncu_potrf.log 2>&1; echo "ncu potrf rc=$?"
