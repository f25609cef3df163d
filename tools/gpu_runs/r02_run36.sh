#!/bin/bash
# round 2, call 36: ncu --set full of the new operand paths: stem (mode 5) + 3x3 64->64 (mode 4) in ResNet-101, DenseNet 1x1 convs with the transform warps
mkdir -p gpurun_out
NC="python bench.py --steps 1 --warmup 3 --masks-per-step 384 --micro-batch 384 --streams 1 --no-cpu-baseline --no-gp --no-library-bar --refine-ties 0"
$NC > gpurun_out/nc_plain.log 2>&1 && timeout 600 ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k 'regex:conv_tc3_kernel<\(int\)64' -s 12 -c 4 -o gpurun_out/r02_prof_tc3_64 $NC > gpurun_out/ncu_tc3_64.log 2>&1; echo "ncu resnet tc3<64> rc=$?"
ND="python bench.py --arch densenet121 --images 1 --masks-per-image 384 --steps 1 --warmup 3 --micro-batch 384 --streams 1 --no-cpu-baseline --no-gp --no-library-bar --refine-ties 0"
$ND > gpurun_out/nd_plain.log 2>&1 && timeout 600 ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k 'regex:conv_tc3_kernel<\(int\)128' -s 360 -c 3 -o gpurun_out/r02_prof_densenet_1x1 $ND > gpurun_out/ncu_dn_1x1.log 2>&1; echo "ncu densenet 1x1 rc=$?"
ls -la gpurun_out/r02_prof_tc3_64.ncu-rep gpurun_out/r02_prof_densenet_1x1.ncu-rep
