#!/bin/bash
timeout 600 python -m pytest tests/test_gpu_classifier.py -q -m gpu -p no:cacheprovider -x -k "fused_expand" 2>&1 | tail -4
timeout 300 python __graft_entry__.py --smoke 2>&1 | tail -2
