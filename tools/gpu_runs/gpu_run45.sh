#!/bin/bash
# closing evidence, v5 = v4 + fused expansion/reduction kernel
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -q -m gpu -p no:cacheprovider > gpurun_out/pytest_gpu_final.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/pytest_gpu_final.log
timeout 900 python bench.py --profile-json gpurun_out/profile_v6_mb384.json > gpurun_out/bench_default_v6.json 2>gpurun_out/bench_default_v6.err; echo "bench rc=$?"; cut -c1-200 gpurun_out/bench_default_v6.json
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_reference_v6.json 2>/dev/null; cut -c1-160 gpurun_out/bench_reference_v6.json
timeout 600 python bench.py --arch densenet121 --no-cpu-baseline --no-gp > gpurun_out/bench_densenet_v6.json 2>&1; cut -c1-120 gpurun_out/bench_densenet_v6.json
NC="python bench.py --steps 1 --warmup 3 --masks-per-step 256 --micro-batch 256 --no-cpu-baseline --no-gp"
$NC > gpurun_out/plain_v6.log 2>&1 || { echo "plain run failed"; tail -5 gpurun_out/plain_v6.log; exit 1; }
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_v6.csv $NC > gpurun_out/ncu_list_v6.log 2>&1
echo "launch list rc=$? lines=$(wc -l < gpurun_out/launches_v6.csv)"; python tools/agg_launches.py gpurun_out/launches_v6.csv 12
timeout 1200 ncu --set full --clock-control none --import-source on -k regex:conv_fused_ca_kernel -s 96 -c 3 -f -o gpurun_out/prof_v6_fused_layer3 $NC > gpurun_out/ncu_full_v6a.log 2>&1; echo "full fused rc=$?"
ls -la gpurun_out/*v6*.ncu-rep | tail -3
