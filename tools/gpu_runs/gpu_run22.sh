#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_classifier.py tests/test_gpu_engine.py tests/test_gpu_masks.py -q -m gpu -p no:cacheprovider -x > gpurun_out/pytest_gpu_v3.log 2>&1
echo "pytest rc=$?"; tail -4 gpurun_out/pytest_gpu_v3.log
run() { name=$1; mb=$2; shift; shift; env "$@" timeout 300 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-gp --micro-batch $mb --masks-per-step $((mb*8)) --profile-json gpurun_out/profile_$name.json > gpurun_out/bench_$name.log 2>gpurun_out/bench_$name.err; echo -n "$name: "; python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/bench_$name.log").read().strip().splitlines()[-1]); r=d["roofline"]
    print(round(d["value"]), "evals/s  e2e", round(d["e2e"]["value"]), " tc TF/s", round(r["achieved"],1), r["per_kind_ms"])
except Exception as e: print("ERR", e, open("gpurun_out/bench_$name.err").read()[-600:])
PY
}
run stem4_mb256 256 A=1
