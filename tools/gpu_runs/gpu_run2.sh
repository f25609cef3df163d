#!/bin/bash
# second GPU call: tests, per-op profile, config sweep, ncu launch list + one full capture of the hot kernel
mkdir -p gpurun_out
timeout 900 python -m pytest tests -q -m gpu -p no:cacheprovider > gpurun_out/pytest_gpu.log 2>&1
echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
tail -4 gpurun_out/pytest_gpu.log
B="python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-gp"
timeout 300 $B --profile-json gpurun_out/profile_mb256.json > gpurun_out/bench_mb256.log 2>gpurun_out/bench_mb256.err
timeout 300 $B --micro-batch 128 --profile-json gpurun_out/profile_mb128.json > gpurun_out/bench_mb128.log 2>&1
timeout 300 $B --micro-batch 512 > gpurun_out/bench_mb512.log 2>&1
timeout 300 $B --graph > gpurun_out/bench_graph.log 2>&1
NIB_TC_BLOCK_N=256 timeout 300 $B --profile-json gpurun_out/profile_bn256.json > gpurun_out/bench_bn256.log 2>&1
NIB_TC_BLOCK_N=64 timeout 300 $B --profile-json gpurun_out/profile_bn64.json > gpurun_out/bench_bn64.log 2>&1
for f in mb256 mb128 mb512 graph bn256 bn64; do echo -n "$f: "; python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/bench_$f.log").read().strip().splitlines()[-1]); r=d["roofline"]
    print(round(d["value"]), "evals/s  e2e", round(d["e2e"]["value"]), " tc TF/s", round(r["achieved"],1), r["per_kind_ms"])
except Exception as e: print("ERR", e)
PY
done
# ncu: launch list of the bench command (small workload so the capture stays short), then the hot kernel
NC="python bench.py --steps 1 --warmup 3 --masks-per-step 256 --micro-batch 256 --no-cpu-baseline --no-gp"
$NC > gpurun_out/plain.log 2>&1 && \
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -s 330 -c 240 --csv --log-file gpurun_out/launches.csv $NC > gpurun_out/ncu_list.log 2>&1
$NC > gpurun_out/plain2.log 2>&1 && \
timeout 1200 ncu --set full --clock-control none --import-source on -k regex:conv_tc_kernel -s 140 -c 6 -o gpurun_out/prof_conv_tc $NC > gpurun_out/ncu_full.log 2>&1
ls -la gpurun_out | tail -20
