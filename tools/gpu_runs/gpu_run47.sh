#!/bin/bash
timeout 600 python -m pytest tests/test_gpu_gp.py tests/test_gpu_ski.py tests/test_gpu_engine.py -q -m gpu -p no:cacheprovider -x 2>&1 | tail -3
echo "--- lookahead"; timeout 200 python tools/gp_once.py 8192 4 2>&1 | tail -2
echo "--- lookahead NBC=256"; NIB_GP_NBC=256 timeout 200 python tools/gp_once.py 8192 3 2>&1 | tail -1
echo "--- no lookahead"; NIB_GP_NO_LOOKAHEAD=1 timeout 200 python tools/gp_once.py 8192 3 2>&1 | tail -1
timeout 300 python tools/gp_profile.py 2>&1 | cut -c1-330 | tail -2
