#!/bin/bash
# ncu evidence for the CTA-pair kernel: launch list of the bench command + full-set capture of layer-3 and layer-1 launches
mkdir -p gpurun_out
NC="python bench.py --steps 1 --warmup 3 --masks-per-step 256 --micro-batch 256 --no-cpu-baseline --no-gp"
$NC > gpurun_out/plain_v3.log 2>&1 || { echo "plain run failed"; tail -5 gpurun_out/plain_v3.log; exit 1; }
tail -c 400 gpurun_out/plain_v3.log
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_v3.csv $NC > gpurun_out/ncu_list_v3.log 2>&1
echo "launch list rc=$? lines=$(wc -l < gpurun_out/launches_v3.csv)"
timeout 1200 ncu --set full --clock-control none --import-source on -k regex:conv_tc3_kernel -s 352 -c 6 -f -o gpurun_out/prof_conv_tc3_layer3 $NC > gpurun_out/ncu_full_v3a.log 2>&1
echo "full layer3 rc=$?"
timeout 1200 ncu --set full --clock-control none --import-source on -k regex:conv_tc3_kernel -s 312 -c 11 -f -o gpurun_out/prof_conv_tc3_layer1 $NC > gpurun_out/ncu_full_v3b.log 2>&1
echo "full layer1 rc=$?"
ls -la gpurun_out/*.ncu-rep
