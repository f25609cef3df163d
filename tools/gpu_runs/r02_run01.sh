#!/bin/bash
# round 2, call 1: baseline pytest -m gpu + diagnostics (margins, fp32 re-score cost, micro-batch/L2 sweep)
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/r02_pytest_baseline.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/r02_pytest_baseline.log
timeout 900 python tools/r02_diag1.py > gpurun_out/r02_diag1.log 2>&1; echo "diag rc=$?"; tail -40 gpurun_out/r02_diag1.log
