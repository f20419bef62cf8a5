#!/bin/bash
mkdir -p gpurun_out
for v in "1 1" "1 2" "2 1"; do set -- $v
  UB200_FPROP_CLUSTER=$1 UB200_FPROP_MSUB=$2 timeout -s KILL 200 python bench.py --steps 10 --warmup 3 --skip-cpu --skip-haar --dump-kernels > gpurun_out/variant_c$1_m$2.log 2>&1
  cp gpurun_out/conv_launch_table.json gpurun_out/conv_table_c$1_m$2.json
  echo "cluster=$1 msub=$2: $(grep -o '"ms_per_step": [0-9.]*' gpurun_out/variant_c$1_m$2.log)"
done
