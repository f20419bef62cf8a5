#!/bin/bash
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv
L=gpurun_out/wgrad_cold2.log; : > $L
r() { UB200_WGRAD_PAIR=$1 UB200_WGRAD_STAGES=$2 PROBE_COLD=$3 PROBE_BIG=0,4,5 timeout 100 python tools/wgrad_probe.py 2>&1 | grep -E "time " | sed "s/^/cold=$3 /" >> $L; }
r 0 6 0; r 2 6 0; r 1 6 0; r 2 8 0; r 2 5 0; r 0 6 1; r 2 6 1; r 0 6 0; r 2 6 0
cut -c1-110 $L
