#!/bin/bash
# round 2, call 32: per-piece timing of one acquisition round (configs[3]) with and without the tie policy
mkdir -p gpurun_out
timeout 300 python tools/r02_bo_round_probe.py > gpurun_out/r02_bo_round_probe.txt 2>&1; echo "rc=$?"; tail -4 gpurun_out/r02_bo_round_probe.txt | cut -c1-300
