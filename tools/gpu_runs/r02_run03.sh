#!/bin/bash
# round 2, call 3: full pytest -m gpu, flip study (16384 masks), VGP tests
mkdir -p gpurun_out
python -m pytest tests -m gpu -q -s > gpurun_out/r02_pytest_3.log 2>&1; echo "pytest rc=$?"; grep -E "passed|failed|FAILED|Error|bench-config|resnet56 bf16" gpurun_out/r02_pytest_3.log | cut -c1-300 | tail -30
timeout 600 python tools/r02_diag2.py 16384 resnet101 > gpurun_out/r02_diag2.log 2>&1; echo "diag2 rc=$?"; tail -60 gpurun_out/r02_diag2.log
