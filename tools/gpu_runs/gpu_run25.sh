#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -q -m gpu -p no:cacheprovider -x > gpurun_out/pytest_gpu_r25.log 2>&1
echo "pytest rc=$?"; tail -4 gpurun_out/pytest_gpu_r25.log
for s in 1 2 3; do
timeout 300 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-gp --streams $s > gpurun_out/bench_streams$s.log 2>gpurun_out/bench_streams$s.err || tail -5 gpurun_out/bench_streams$s.err
python - <<PY
import json
d=json.loads(open("gpurun_out/bench_streams$s.log").read().strip().splitlines()[-1])
print("streams $s:", round(d["value"]), "e2e", round(d["e2e"]["value"]), "launches", d["gpu_launches"], d["clocks"])
PY
done
NIB_TC_SERP=1 timeout 300 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-gp --streams 2 > gpurun_out/bench_streams2_serp.log 2>&1; python -c "
import json; d=json.loads(open('gpurun_out/bench_streams2_serp.log').read().strip().splitlines()[-1]); print('serp streams 2:', round(d['value']))"
