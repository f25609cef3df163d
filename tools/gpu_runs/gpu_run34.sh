#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_classifier.py tests/test_gpu_engine.py -q -m gpu -p no:cacheprovider -x -k "resnet56 or cifar or stream" 2>&1 | tail -15
timeout 600 python tools/bench_configs.py 2>/dev/null | grep "configs\[1\]" | cut -c1-300
