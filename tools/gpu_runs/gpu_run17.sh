#!/bin/bash
# per-launch DRAM traffic + tensor-pipe activity of every conv launch of one ResNet-101 forward (micro-batch 256)
mkdir -p gpurun_out
NC="python bench.py --steps 1 --warmup 3 --masks-per-step 256 --micro-batch 256 --no-cpu-baseline --no-gp"
$NC > gpurun_out/plain_v3b.log 2>&1 || { echo "plain run failed"; tail -5 gpurun_out/plain_v3b.log; exit 1; }
timeout 1500 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed,sm__cycles_active.avg,sm__cycles_elapsed.max,lts__t_sector_hit_rate.pct,l1tex__m_xbar2l1tex_read_bytes.sum \
  --clock-control none -k regex:"conv_tc|pool_|fc_kernel|mask_synth|halo_zero|score_kernel" -s 330 -c 112 --csv --log-file gpurun_out/ncu_all_layers_v3.csv $NC > gpurun_out/ncu_all_layers_v3.log 2>&1
echo "rc=$? lines=$(wc -l < gpurun_out/ncu_all_layers_v3.csv)"
