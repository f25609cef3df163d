#!/bin/bash
timeout 120 python tools/fused_dbg.py 256 2>&1 | grep -A8 "fused K1" | tail -9
timeout 300 python -m pytest tests/test_gpu_classifier.py -q -m gpu -p no:cacheprovider -x -k "fused_expand or resnet101_bf16 or resnet56" 2>&1 | tail -3
