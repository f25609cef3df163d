#!/bin/bash
timeout 900 python -m pytest tests/test_gpu_ski.py -q -m gpu -p no:cacheprovider -x 2>&1 | tail -25
