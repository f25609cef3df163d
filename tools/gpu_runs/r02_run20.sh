#!/bin/bash
# round 2, call 20: ncu pass over every launch of one 384-mask forward (shipped state) + ncu --set full of the stem / 3x3 64 kernel
mkdir -p gpurun_out
NC="python bench.py --steps 1 --warmup 3 --masks-per-step 384 --micro-batch 384 --streams 1 --no-cpu-baseline --no-gp --no-library-bar --refine-ties 0"
M="gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed,sm__cycles_active.avg,sm__cycles_elapsed.max,lts__t_sector_hit_rate.pct,l1tex__m_xbar2l1tex_read_bytes.sum"
$NC > gpurun_out/nc_plain.log 2>&1 && timeout 900 ncu --metrics $M --clock-control none --csv --page raw --log-file gpurun_out/r02_all_launches_raw.csv -s 250 -c 200 $NC > gpurun_out/ncu_all.log 2>&1; echo "ncu all rc=$? lines=$(wc -l < gpurun_out/r02_all_launches_raw.csv)"
python tools/ncu_all_launches.py gpurun_out/r02_all_launches_raw.csv gpurun_out/r02_ncu_all_launches_one_forward.csv
$NC > gpurun_out/nc_plain2.log 2>&1 && timeout 900 ncu --set full --clock-control none --import-source on -k regex:"conv_tc3_kernel<64" -s 12 -c 4 -o gpurun_out/r02_prof_tc3_64 $NC > gpurun_out/ncu_tc3_64.log 2>&1; echo "ncu tc3<64> rc=$?"
