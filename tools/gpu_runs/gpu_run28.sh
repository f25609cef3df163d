#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_gp.py tests/test_gpu_engine.py -q -m gpu -p no:cacheprovider -x > gpurun_out/pytest_gp_v2.log 2>&1
echo "pytest rc=$?"; tail -5 gpurun_out/pytest_gp_v2.log
echo "--- v2"; timeout 200 python tools/gp_once.py 8192 3 2>&1 | tail -2
echo "--- v2 NBC=512"; NIB_GP_NBC=512 timeout 200 python tools/gp_once.py 8192 3 2>&1 | tail -1
echo "--- v2 NBC=128"; NIB_GP_NBC=128 timeout 200 python tools/gp_once.py 8192 3 2>&1 | tail -1
echo "--- v1"; NIB_GP_V1=1 timeout 200 python tools/gp_once.py 8192 3 2>&1 | tail -1
timeout 300 python tools/gp_profile.py 2>&1 | tail -3
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/gp_launches_v2.csv python tools/gp_once.py 8192 1 > gpurun_out/gp_ncu.log 2>&1
python tools/agg_launches.py gpurun_out/gp_launches_v2.csv
