#!/bin/bash
# round 2, call 28: BN-ReLU applied to the A tiles inside the pair kernel (no pack pass): parity, DenseNet bench A/B, ResNet bench unchanged?
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_classifier.py tests/test_gpu_engine.py tests/test_gpu_bench_config.py -q > gpurun_out/r02_pytest_28.log 2>&1; echo "pytest rc=$?"; tail -6 gpurun_out/r02_pytest_28.log
timeout 600 python bench.py --arch densenet121 --images 2 --masks-per-image 3072 --no-cpu-baseline --no-library-bar --no-gp --refine-ties 0 --profile-json gpurun_out/r02_per_op_profile_densenet121.json > gpurun_out/r02_bench_28_densenet.json 2> gpurun_out/r02_bench_28.err; echo "bench rc=$?"; cut -c1-200 gpurun_out/r02_bench_28_densenet.json
NIB_TC_NO_XFORM=1 timeout 600 python bench.py --arch densenet121 --images 2 --masks-per-image 3072 --no-cpu-baseline --no-library-bar --no-gp --refine-ties 0 > gpurun_out/r02_bench_28_densenet_noxform.json 2>> gpurun_out/r02_bench_28.err; cut -c1-200 gpurun_out/r02_bench_28_densenet_noxform.json
python tools/prof_table.py gpurun_out/r02_per_op_profile_densenet121.json 2>/dev/null | head -8
timeout 600 python bench.py --no-cpu-baseline --no-library-bar --no-gp > gpurun_out/r02_bench_28.json 2>> gpurun_out/r02_bench_28.err; cut -c1-200 gpurun_out/r02_bench_28.json
tail -5 gpurun_out/r02_bench_28.err
