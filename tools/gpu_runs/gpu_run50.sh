#!/bin/bash
# final sanity: full GPU test suite, smoke, micro-batch sweep of the final kernels
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -q -m gpu -p no:cacheprovider > gpurun_out/pytest_gpu_final.log 2>&1; echo "pytest rc=$?"; tail -2 gpurun_out/pytest_gpu_final.log
timeout 300 python __graft_entry__.py --smoke 2>&1 | tail -1
run() { name=$1; mb=$2; shift; shift; timeout 300 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-gp --micro-batch $mb --masks-per-step $((mb*8)) "$@" > gpurun_out/bench_$name.log 2>gpurun_out/bench_$name.err; echo -n "$name: "; python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/bench_$name.log").read().strip().splitlines()[-1]); r=d["roofline"]
    print(round(d["value"]), "evals/s  e2e", round(d["e2e"]["value"]), " iso TF/s", round(r["achieved"],1), "in-step", round(r["achieved_in_step"],1), d["clocks"]["sm_mhz"])
except Exception as e: print("ERR", e, open("gpurun_out/bench_$name.err").read()[-600:])
PY
}
run f_mb320 320
run f_mb384 384
run f_mb448 448
run f_mb512 512
run f_mb384_s3 384 --streams 3
