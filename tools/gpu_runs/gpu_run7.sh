#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_classifier.py tests/test_gpu_engine.py -q -m gpu -p no:cacheprovider > gpurun_out/pytest_gpu.log 2>&1
echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
tail -4 gpurun_out/pytest_gpu.log
B="python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-gp"
run() { name=$1; shift; env "$@" timeout 300 $B --profile-json gpurun_out/profile_$name.json > gpurun_out/bench_$name.log 2>gpurun_out/bench_$name.err; echo -n "$name: "; python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/bench_$name.log").read().strip().splitlines()[-1]); r=d["roofline"]
    print(round(d["value"]), "evals/s  e2e", round(d["e2e"]["value"]), " tc TF/s", round(r["achieved"],1), r["per_kind_ms"])
except Exception as e: print("ERR", e, open("gpurun_out/bench_$name.err").read()[-400:])
PY
}
run pf0 NIB_TC_PF=0
run pf1 NIB_TC_PF=1
run pf2 NIB_TC_PF=2
run pf4 NIB_TC_PF=4
B="python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-gp --micro-batch 512"
run mb512_pf2 NIB_TC_PF=2
