#!/bin/bash
# round 2, call 26: resident-patch mode generalised to Cin 128 / 32 outputs (DenseNet growth convs): parity, DenseNet nets, bench
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_classifier.py -q > gpurun_out/r02_pytest_26.log 2>&1; echo "classifier pytest rc=$?"; tail -6 gpurun_out/r02_pytest_26.log
timeout 600 python bench.py --arch densenet121 --images 2 --masks-per-image 3072 --no-cpu-baseline --no-library-bar --no-gp --refine-ties 0 --profile-json gpurun_out/r02_per_op_profile_densenet121_26.json > gpurun_out/r02_bench_26_densenet.json 2> gpurun_out/r02_bench_26.err; echo "bench rc=$?"; cut -c1-200 gpurun_out/r02_bench_26_densenet.json
NIB_TC_NO_HALO=1 timeout 600 python bench.py --arch densenet121 --images 2 --masks-per-image 3072 --no-cpu-baseline --no-library-bar --no-gp --refine-ties 0 > gpurun_out/r02_bench_26_densenet_nohalo.json 2>> gpurun_out/r02_bench_26.err; cut -c1-200 gpurun_out/r02_bench_26_densenet_nohalo.json
python tools/prof_table.py gpurun_out/r02_per_op_profile_densenet121_26.json 2>/dev/null | head -8
timeout 600 python bench.py --no-cpu-baseline --no-library-bar --no-gp > gpurun_out/r02_bench_26.json 2>> gpurun_out/r02_bench_26.err; cut -c1-200 gpurun_out/r02_bench_26.json
