#!/bin/bash
# round 2, call 33 (last): drop-in scripts end to end (CIFAR on the x3 default, ImageNet), configs[0/1/3] with warm-up, full pytest -m gpu
mkdir -p gpurun_out
cd /root/repo
timeout 300 python generate_gp_training_data_cifar.py --num_mask_samples 512 --no-write > gpurun_out/r02_dropin_cifar.log 2>&1; echo "cifar drop-in rc=$?"; tail -3 gpurun_out/r02_dropin_cifar.log | cut -c1-200
timeout 300 python generate_gp_training_data_cifar.py --num_mask_samples 512 --no-write --precision fp32 > gpurun_out/r02_dropin_cifar_fp32.log 2>&1; echo "cifar fp32 rc=$?"; tail -2 gpurun_out/r02_dropin_cifar_fp32.log | cut -c1-200
timeout 600 python tools/bench_configs.py > gpurun_out/r02_bench_configs_0_1_3.jsonl 2> gpurun_out/r02_bench_configs.err; echo "configs rc=$?"; cut -c1-300 gpurun_out/r02_bench_configs_0_1_3.jsonl; tail -2 gpurun_out/r02_bench_configs.err
python -m pytest tests -m gpu -q > gpurun_out/r02_pytest_33.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/r02_pytest_33.log
