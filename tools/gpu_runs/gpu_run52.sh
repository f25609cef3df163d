#!/bin/bash
# ncu launch list of the DEFAULT bench command (first 2000 launches: three full steps of 8 micro-batches of 384 masks)
mkdir -p gpurun_out
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 2000 --csv --log-file gpurun_out/launches_default_v6.csv python bench.py --no-cpu-baseline > gpurun_out/ncu_default_v6.log 2>&1
echo "rc=$? lines=$(wc -l < gpurun_out/launches_default_v6.csv)"; python tools/agg_launches.py gpurun_out/launches_default_v6.csv 12
