#!/bin/bash
# warm-cache ncu launch list of one eager train step (after the same command exited 0 without ncu)
# usage: tools/gpu_launchlist.sh <tag> [skip] [count]
tag=${1:-x}; skip=${2:-3000}; cnt=${3:-1100}
mkdir -p gpurun_out
timeout -s KILL 300 python bench.py --steps 1 --warmup 3 --no-graph --skip-cpu --skip-haar > gpurun_out/plain.log 2>&1 || { echo plain failed; tail -5 gpurun_out/plain.log; exit 1; }
timeout -s KILL 800 ncu --metrics gpu__time_duration.sum --clock-control none --cache-control none --profile-from-start off --csv \
  --log-file gpurun_out/launches_$tag.csv python bench.py --steps 1 --warmup 3 --no-graph --skip-cpu --skip-haar --profile-step > gpurun_out/ncu.log 2>&1
echo ncu exit=$?
python tools/summarize_launches.py gpurun_out/launches_$tag.csv 60 | cut -c1-220
