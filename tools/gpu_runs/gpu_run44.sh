#!/bin/bash
timeout 120 python tools/fused_dbg.py 256 2>&1 | grep -A8 "fused K1" | tail -9
timeout 300 python -m pytest tests/test_gpu_classifier.py tests/test_gpu_engine.py -q -m gpu -p no:cacheprovider -x 2>&1 | tail -3
