#!/bin/bash
# warm ncu launch lists of one eager step of the other BASELINE configs (each after the same command exited 0 without ncu)
mkdir -p gpurun_out
for c in c3 c4 c5 c1; do
  CMD="python bench.py --config $c --steps 1 --warmup 3 --no-graph --skip-cpu --skip-lib --skip-haar --profile-step"
  timeout -s KILL 300 $CMD > gpurun_out/plain_$c.log 2>&1 || { echo "plain $c failed"; tail -5 gpurun_out/plain_$c.log; continue; }
  timeout -s KILL 600 ncu --metrics gpu__time_duration.sum --clock-control none --cache-control none --profile-from-start off --csv \
    --log-file gpurun_out/launches_$c.csv $CMD > gpurun_out/ncu_$c.log 2>&1
  echo "ncu $c exit=$?"
  python tools/summarize_launches.py gpurun_out/launches_$c.csv 22 | cut -c1-200
done
