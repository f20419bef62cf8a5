#!/bin/bash
# ncu --set full captures of the top kernels of one eager config-2 step (after the same command exited 0 without ncu),
# plus the warm launch list of one config-3 step
mkdir -p gpurun_out
CMD="python bench.py --steps 1 --warmup 3 --no-graph --skip-cpu --skip-haar --skip-lib --profile-step"
timeout -s KILL 300 $CMD > gpurun_out/plain.log 2>&1 || { echo plain failed; tail -5 gpurun_out/plain.log; exit 1; }
for k in conv_fprop_kernel conv_wgrad_kernel bgemm256_kernel; do
  timeout -s KILL 600 ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:$k -c 12 \
     -o gpurun_out/r02_prof_$k -f $CMD > gpurun_out/ncu_$k.log 2>&1
  echo "ncu $k exit=$?"
done
CMD3="python bench.py --config c3 --steps 1 --warmup 3 --no-graph --skip-cpu --skip-lib --profile-step"
timeout -s KILL 300 $CMD3 > gpurun_out/plain_c3.log 2>&1 && \
timeout -s KILL 600 ncu --metrics gpu__time_duration.sum --clock-control none --cache-control none --profile-from-start off --csv \
  --log-file gpurun_out/launches_c3.csv $CMD3 > gpurun_out/ncu_c3.log 2>&1
echo "ncu c3 exit=$?"
python tools/summarize_launches.py gpurun_out/launches_c3.csv 30 | cut -c1-200
ls -la gpurun_out/*.ncu-rep
