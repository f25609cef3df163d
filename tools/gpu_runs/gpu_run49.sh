#!/bin/bash
mkdir -p gpurun_out
NC="python bench.py --steps 1 --warmup 3 --masks-per-step 256 --micro-batch 256 --streams 1 --no-cpu-baseline --no-gp"
M="gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed,sm__cycles_active.avg,sm__cycles_elapsed.max,lts__t_sector_hit_rate.pct,l1tex__m_xbar2l1tex_read_bytes.sum"
timeout 900 ncu --metrics $M --clock-control none --csv --page raw --log-file gpurun_out/all_launches_v6_raw.csv -s 250 -c 200 $NC > gpurun_out/ncu_all_v6.log 2>&1; echo "rc=$? lines=$(wc -l < gpurun_out/all_launches_v6_raw.csv)"
python tools/ncu_all_launches.py gpurun_out/all_launches_v6_raw.csv gpurun_out/all_launches_v6.csv
