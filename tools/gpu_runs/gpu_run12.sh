#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_classifier.py tests/test_gpu_engine.py -q -m gpu -p no:cacheprovider -x > gpurun_out/pytest_gpu_v3.log 2>&1
echo "pytest rc=$?"; tail -3 gpurun_out/pytest_gpu_v3.log
timeout 600 python tools/tc_diag.py > gpurun_out/tc_diag_v3.log 2>&1; grep -c '"n_bad": 0' gpurun_out/tc_diag.txt; grep -v '"n_bad": 0' gpurun_out/tc_diag.txt | cut -c1-250
run() { name=$1; mb=$2; shift; shift; env "$@" timeout 300 python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-gp --micro-batch $mb --masks-per-step $((mb*8)) --profile-json gpurun_out/profile_$name.json > gpurun_out/bench_$name.log 2>gpurun_out/bench_$name.err; echo -n "$name: "; python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/bench_$name.log").read().strip().splitlines()[-1]); r=d["roofline"]
    print(round(d["value"]), "evals/s  e2e", round(d["e2e"]["value"]), " tc TF/s", round(r["achieved"],1), r["per_kind_ms"])
except Exception as e: print("ERR", e, open("gpurun_out/bench_$name.err").read()[-600:])
PY
}
run v3d_mb256 256 NIB_TC_VER=3
run v3d_mb288 288 NIB_TC_VER=3
run v3d_mb512 512 NIB_TC_VER=3
NIB_TC_DBG=1 timeout 600 python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-gp --micro-batch 256 --masks-per-step 256 > gpurun_out/dbg_bench.log 2> gpurun_out/dbg_roles.txt
grep "^\[tc3" gpurun_out/dbg_roles.txt | tail -208 > gpurun_out/dbg_roles_last.txt
