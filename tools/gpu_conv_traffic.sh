#!/bin/bash
# DRAM / L2->SM traffic, duration and tensor-pipe activity of EVERY conv launch of one eager config-2 step (metrics pass, csv)
mkdir -p gpurun_out
CMD="python bench.py --steps 1 --warmup 3 --no-graph --skip-cpu --skip-haar --skip-lib --profile-step"
timeout -s KILL 300 $CMD > gpurun_out/plain.log 2>&1 || { echo plain failed; tail -5 gpurun_out/plain.log; exit 1; }
M=gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,l1tex__m_xbar2l1tex_read_bytes.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active,launch__grid_size
timeout -s KILL 900 ncu --metrics $M --clock-control none --profile-from-start off -k regex:'conv_fprop_kernel|conv_wgrad_kernel' --csv \
  --log-file gpurun_out/conv_traffic.csv $CMD > gpurun_out/ncu_traffic.log 2>&1
echo "ncu exit=$?"; wc -l gpurun_out/conv_traffic.csv
