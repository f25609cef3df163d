#!/bin/bash
# round-1 closing evidence: full GPU test log, default bench line (+ reference arm), ncu launch list and full-set captures
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -q -m gpu -p no:cacheprovider > gpurun_out/pytest_gpu_final.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/pytest_gpu_final.log
timeout 900 python bench.py > gpurun_out/bench_default_v4.json 2>gpurun_out/bench_default_v4.err; echo "bench rc=$?"; cut -c1-300 gpurun_out/bench_default_v4.json
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_reference_v4.json 2>gpurun_out/bench_reference_v4.err; echo "ref rc=$?"; cut -c1-300 gpurun_out/bench_reference_v4.json
timeout 600 python bench.py --arch densenet121 --no-cpu-baseline --no-gp > gpurun_out/bench_densenet_v4.json 2>&1; cut -c1-200 gpurun_out/bench_densenet_v4.json
NC="python bench.py --steps 1 --warmup 3 --masks-per-step 256 --micro-batch 256 --no-cpu-baseline --no-gp"
$NC > gpurun_out/plain_v4.log 2>&1 || { echo "plain run failed"; tail -5 gpurun_out/plain_v4.log; exit 1; }
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_v4.csv $NC > gpurun_out/ncu_list_v4.log 2>&1
echo "launch list rc=$? lines=$(wc -l < gpurun_out/launches_v4.csv)"; python tools/agg_launches.py gpurun_out/launches_v4.csv 10
timeout 1200 ncu --set full --clock-control none --import-source on -k regex:conv_tc3_kernel -s 352 -c 6 -f -o gpurun_out/prof_v4_layer3 $NC > gpurun_out/ncu_full_v4a.log 2>&1; echo "full layer3 rc=$?"
timeout 1200 ncu --set full --clock-control none --import-source on -k regex:conv_tc3_kernel -s 312 -c 5 -f -o gpurun_out/prof_v4_stem_layer1 $NC > gpurun_out/ncu_full_v4b.log 2>&1; echo "full stem rc=$?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:dgemm_sub_kernel -s 40 -c 3 -f -o gpurun_out/prof_v4_dgemm python tools/gp_once.py 8192 1 > gpurun_out/ncu_full_v4c.log 2>&1; echo "full dgemm rc=$?"
ls -la gpurun_out/*.ncu-rep | tail -5
