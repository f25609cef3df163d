#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_classifier.py tests/test_gpu_engine.py -q -m gpu -p no:cacheprovider -x 2>&1 | tail -6
run() { name=$1; mb=$2; shift; shift; env "$@" timeout 300 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-gp --micro-batch $mb --masks-per-step $((mb*8)) --profile-json gpurun_out/profile_$name.json > gpurun_out/bench_$name.log 2>gpurun_out/bench_$name.err; echo -n "$name: "; python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/bench_$name.log").read().strip().splitlines()[-1]); r=d["roofline"]
    print(round(d["value"]), "evals/s  e2e", round(d["e2e"]["value"]), " iso TF/s", round(r["achieved"],1), "in-step", round(r["achieved_in_step"],1), d["clocks"]["sm_mhz"], r["per_kind_ms"])
except Exception as e: print("ERR", e, open("gpurun_out/bench_$name.err").read()[-600:])
PY
}
run fuse1_mb256 256 NIB_TC_FUSE=1
run fuse0_mb256 256 NIB_TC_FUSE=0
run fuse1_mb384 384 NIB_TC_FUSE=1
run fuse0_mb384 384 NIB_TC_FUSE=0
python - <<'PY'
import json,collections
for n in ['fuse1_mb256','fuse0_mb256']:
    d=json.load(open(f'gpurun_out/profile_{n}.json'))
    g=collections.OrderedDict()
    for o in d['ops']:
        if o['kind']!='tc_conv' or o['k']!=1: continue
        k=(o['Hout'],o['Cin'],o['Cout'],o['residual'])
        a=g.setdefault(k,[0,0]); a[0]+=1;a[1]+=o['ms']
    print(n, {k:round(1000*a[1]/a[0],1) for k,a in g.items() if k[3]==1 or k[1]>k[2]})
PY
