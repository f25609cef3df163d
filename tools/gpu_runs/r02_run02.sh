#!/bin/bash
# round 2, call 2: full pytest -m gpu with the tie policy / stream scratch / NCCL C ABI changes, default bench + tie policy off,
# reference arm, first compute-sanitizer passes
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q -s > gpurun_out/r02_pytest_2.log 2>&1; echo "pytest rc=$?"; tail -15 gpurun_out/r02_pytest_2.log | cut -c1-400
timeout 600 python bench.py > gpurun_out/r02_bench_a.json 2> gpurun_out/r02_bench_a.err; echo "bench rc=$?"; cut -c1-1500 gpurun_out/r02_bench_a.json; tail -5 gpurun_out/r02_bench_a.err
timeout 300 python bench.py --refine-ties 0 --no-cpu-baseline --no-gp --no-library-bar > gpurun_out/r02_bench_a_noties.json 2>> gpurun_out/r02_bench_a.err; echo "bench noties rc=$?"; cut -c1-600 gpurun_out/r02_bench_a_noties.json
timeout 300 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r02_bench_a_ref.json 2>> gpurun_out/r02_bench_a.err; cut -c1-300 gpurun_out/r02_bench_a_ref.json
timeout 400 compute-sanitizer --tool memcheck python tools/sanitize_targets.py masks score gp > gpurun_out/r02_sanitizer_memcheck_masks_score_gp.log 2>&1; echo "memcheck rc=$?"; tail -4 gpurun_out/r02_sanitizer_memcheck_masks_score_gp.log
timeout 400 compute-sanitizer --tool racecheck python tools/sanitize_targets.py masks score > gpurun_out/r02_sanitizer_racecheck_masks_score.log 2>&1; echo "racecheck rc=$?"; tail -4 gpurun_out/r02_sanitizer_racecheck_masks_score.log
