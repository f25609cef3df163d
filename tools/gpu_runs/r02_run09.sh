#!/bin/bash
# round 2, call 9: conv_x3_kernel per-layer timing at the tie policy's batch sizes + ncu --set full of its hottest shapes
mkdir -p gpurun_out
python tools/r02_x3_probe.py 80 x3 > gpurun_out/r02_x3_probe_80.log 2>&1; echo "probe80 rc=$?"; cat gpurun_out/r02_x3_probe_80.log | tail -16
python tools/r02_x3_probe.py 256 x3 > gpurun_out/r02_x3_probe_256.log 2>&1; echo "probe256 rc=$?"; cat gpurun_out/r02_x3_probe_256.log | tail -16
NC="python tools/r02_x3_probe.py 80 x3"
$NC > gpurun_out/nc_plain.log 2>&1 && timeout 900 ncu --set full --clock-control none --import-source on -k regex:conv_x3_kernel -s 330 -c 12 -o gpurun_out/r02_prof_x3 $NC > gpurun_out/ncu_x3.log 2>&1; echo "ncu x3 rc=$?"; ls -la gpurun_out/r02_prof_x3.ncu-rep
