#!/bin/bash
# round 2, call 23: micro-batch sweep of the default bench with the shipped kernels (one box)
mkdir -p gpurun_out
for mb in 384 256 512 768 384; do
  timeout 300 python bench.py --micro-batch $mb --no-cpu-baseline --no-library-bar --no-gp > gpurun_out/r02_sweep_mb$mb.json 2> gpurun_out/r02_sweep.err
  python -c "
import json,sys; d=json.loads(open('gpurun_out/r02_sweep_mb$mb.json').read().strip().splitlines()[-1]); print('mb', $mb, round(d['value']), round(d['e2e']['value']), d['ms_per_step'], d['clocks']['sm_mhz'], d['roofline']['frac'])"
done
timeout 300 python bench.py --micro-batch 512 --masks-per-step 4096 --no-cpu-baseline --no-library-bar --no-gp > gpurun_out/r02_sweep_mb512_4096.json 2>> gpurun_out/r02_sweep.err
python -c "
import json,sys; d=json.loads(open('gpurun_out/r02_sweep_mb512_4096.json').read().strip().splitlines()[-1]); print('mb 512 step 4096', round(d['value']), d['ms_per_step'], d.get('near_ties_per_step'))"
