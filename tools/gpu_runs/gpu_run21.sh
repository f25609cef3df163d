#!/bin/bash
# 8-GPU box: bench at N = 8 and 4 (the driver's scaling run repeats 1/2/4/8 at round end)
mkdir -p gpurun_out
nvidia-smi -L | wc -l
for n in 8 4; do
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 2951$n bench.py --gpus $n --steps 5 --warmup 3 > gpurun_out/bench_n$n.log 2> gpurun_out/bench_n$n.err
  echo "bench n$n rc=$?"; python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/bench_n$n.log").read().strip().splitlines()[-1])
    print($n, round(d["value"]), "evals/s e2e", round(d["e2e"]["value"]), "ms/step", round(d["ms_per_step"],2), d["clocks"])
except Exception as e: print("ERR", e, open("gpurun_out/bench_n$n.err").read()[-800:])
PY
done
