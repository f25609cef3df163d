#!/bin/bash
mkdir -p gpurun_out
timeout 600 python tools/tc_diag.py > gpurun_out/tc_diag.log 2>&1
timeout 900 python -m pytest tests -q -m gpu -p no:cacheprovider > gpurun_out/pytest_gpu.log 2>&1
echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
tail -6 gpurun_out/pytest_gpu.log
B="python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-gp"
timeout 300 $B --profile-json gpurun_out/profile_v2_mb256.json > gpurun_out/bench_v2_mb256.log 2>gpurun_out/bench_v2_mb256.err
timeout 300 $B --micro-batch 128 --profile-json gpurun_out/profile_v2_mb128.json > gpurun_out/bench_v2_mb128.log 2>&1
timeout 300 $B --micro-batch 512 > gpurun_out/bench_v2_mb512.log 2>&1
timeout 300 $B --graph > gpurun_out/bench_v2_graph.log 2>&1
NIB_TC_V1=1 timeout 300 $B > gpurun_out/bench_v1again.log 2>&1
NIB_TC_BLOCK_N=64 timeout 300 $B --profile-json gpurun_out/profile_v2_bn64.json > gpurun_out/bench_v2_bn64.log 2>&1
for f in v2_mb256 v2_mb128 v2_mb512 v2_graph v1again v2_bn64; do echo -n "$f: "; python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/bench_$f.log").read().strip().splitlines()[-1]); r=d["roofline"]
    print(round(d["value"]), "evals/s  e2e", round(d["e2e"]["value"]), " tc TF/s", round(r["achieved"],1), r["per_kind_ms"])
except Exception as e: print("ERR", e)
PY
done
tail -3 gpurun_out/bench_v2_mb256.err
