#!/bin/bash
mkdir -p gpurun_out
timeout 200 python tools/gp_once.py 8192 3 > gpurun_out/gp_once.log 2>&1; cat gpurun_out/gp_once.log
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/gp_launches.csv python tools/gp_once.py 8192 1 > gpurun_out/gp_ncu.log 2>&1
python tools/agg_launches.py gpurun_out/gp_launches.csv
