#!/bin/bash
mkdir -p gpurun_out
echo "--- tile 64"; timeout 300 python tools/gp_profile.py 2>&1 | cut -c1-400
timeout 200 python tools/gp_once.py 8192 3 2>&1 | tail -1
echo "--- tile 128"; NIB_GP_TILE_M=128 timeout 300 python tools/gp_profile.py 2>&1 | head -4 | cut -c1-200
NIB_GP_TILE_M=128 timeout 200 python tools/gp_once.py 8192 3 2>&1 | tail -1
timeout 600 python -m pytest tests/test_gpu_gp.py -q -m gpu -p no:cacheprovider -x 2>&1 | tail -2
