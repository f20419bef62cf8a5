#!/bin/bash
# what does the split-K atomic epilogue of conv_wgrad cost?  (UB200_WGRAD_DEBUG: 0 normal, 1 no atomics, 2 no TMEM loads either)
mkdir -p gpurun_out
L=gpurun_out/wgrad_epi.log; : > $L
for d in 0 1 2 3; do
  UB200_WGRAD_DEBUG=$d PROBE_TIME=1 timeout 200 python tools/wgrad_probe.py 2>&1 | grep -E "time|max active|worst" | sed "s/^/DEBUG=$d /" >> $L
done
cat $L
