#!/bin/bash
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_classifier.py -q -m gpu -p no:cacheprovider -x -k "fused_expand" 2>&1 | tail -25
