#!/bin/bash
# is conv_wgrad on DRAM-cold operands (the in-step case) bound by something the CTA pairs do not touch?
mkdir -p gpurun_out
L=gpurun_out/wgrad_cold.log; : > $L
for cfg in "0 6" "2 6" "2 4" "2 8"; do
  set -- $cfg
  UB200_WGRAD_PAIR=$1 UB200_WGRAD_STAGES=$2 PROBE_COLD=1 timeout 200 python tools/wgrad_probe.py 2>&1 | grep -E "time +(32x32|16x16)|worst" | sed "s/^/COLD PAIR=$1 STAGES=$2 /" >> $L
done
UB200_WGRAD_PAIR=2 UB200_WGRAD_STAGES=8 timeout 200 python tools/wgrad_probe.py 2>&1 | grep -E "time +(32x32|16x16)" | sed "s/^/WARM PAIR=2 STAGES=8 /" >> $L
cat $L
