#!/bin/bash
mkdir -p gpurun_out
run() { name=$1; mb=$2; shift; shift; timeout 300 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-gp --micro-batch $mb --masks-per-step $((mb*8)) "$@" > gpurun_out/bench_$name.log 2>gpurun_out/bench_$name.err; echo -n "$name: "; python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/bench_$name.log").read().strip().splitlines()[-1]); r=d["roofline"]
    print(round(d["value"]), "evals/s  e2e", round(d["e2e"]["value"]), " iso TF/s", round(r["achieved"],1), "in-step", round(r["achieved_in_step"],1), d["clocks"]["sm_mhz"])
except Exception as e: print("ERR", e, open("gpurun_out/bench_$name.err").read()[-600:])
PY
}
run mb256_s2 256 --streams 2
run mb288_s2 288 --streams 2
run mb384_s2 384 --streams 2
run mb512_s2 512 --streams 2
run mb512_s1 512 --streams 1
run mb256_s2_graph 256 --streams 2 --graph
run mb192_s3 192 --streams 3
