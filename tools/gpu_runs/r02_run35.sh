#!/bin/bash
# round 2, call 35: NIB_PREC_X1 (fp32 activations, bf16 operands): ResNet-56 parity, engine tie policy, configs[1] timing
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_classifier.py tests/test_gpu_engine.py -q -s -k "x1 or x3_mode or resnet56" > gpurun_out/r02_pytest_35.log 2>&1; echo "pytest rc=$?"; grep -E "resnet56|cifar x1|passed|failed|Error|assert" gpurun_out/r02_pytest_35.log | cut -c1-300 | tail -12
timeout 600 python tools/bench_configs.py > gpurun_out/r02_bench_configs_0_1_3.jsonl 2> gpurun_out/r02_bench_configs.err; echo "configs rc=$?"; grep "configs\[1\]" gpurun_out/r02_bench_configs_0_1_3.jsonl | cut -c1-300; tail -2 gpurun_out/r02_bench_configs.err
