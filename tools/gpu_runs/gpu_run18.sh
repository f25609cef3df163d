#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_gp.py -q -m gpu -p no:cacheprovider -x > gpurun_out/pytest_gpu_gp.log 2>&1
echo "pytest rc=$?"; tail -3 gpurun_out/pytest_gpu_gp.log
python tools/gp_profile.py; python - <<"PY"
import json, numpy as np, torch, sys
sys.path.insert(0, '.')
import bench, network_interpretation_imagenet_b200 as nib
for n in (1024, 4096, 8192):
    g = bench.gp_bench(nib, torch, np, n, 50)
    print(json.dumps(g))
PY
