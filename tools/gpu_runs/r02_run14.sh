#!/bin/bash
# round 2, call 14: ncu --set full of the fp64 tensor-core GEMM at K = 4096 (launch 12 of tools/gp_profile.py) and at K = 512
mkdir -p gpurun_out
python tools/gp_profile.py > /dev/null 2>&1 && timeout 600 ncu --set full --clock-control none --import-source on -k regex:dgemm_sub_kernel -s 11 -c 1 -o gpurun_out/r02_dgemm_k4096 python tools/gp_profile.py > gpurun_out/ncu_dgemm.log 2>&1; echo "ncu dgemm rc=$?"
