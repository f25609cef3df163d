#!/bin/bash
mkdir -p gpurun_out
NIB_TC_DBG=1 timeout 600 python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-gp --micro-batch 256 --masks-per-step 256 > gpurun_out/dbg_bench.log 2> gpurun_out/dbg_roles.txt
grep "^\[tc3" gpurun_out/dbg_roles.txt | tail -208 > gpurun_out/dbg_roles_last.txt
wc -l gpurun_out/dbg_roles_last.txt
