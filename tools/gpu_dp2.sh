#!/bin/bash
# 2-GPU data parallel: default (graph fwd+bwd, then one all-reduce + tail) vs NCCL captured inside the graph (bucketed, overlapped)
mkdir -p gpurun_out
N=${N:-2}
for mode in 0 1 0 1; do
  UB200_DP_GRAPH_NCCL=$mode timeout -s KILL 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $((29500 + RANDOM % 500)) \
     bench.py --gpus $N --steps 30 --warmup 5 --skip-cpu --skip-haar --skip-lib > gpurun_out/dp${N}_nccl$mode.log 2>&1
  echo "DP_GRAPH_NCCL=$mode exit=$?"
  tail -1 gpurun_out/dp${N}_nccl$mode.log | python -c "import json,sys; d=json.loads(sys.stdin.read()); print(round(d['ms_per_step'],3),'ms/step', round(d['value']),'img/s', d.get('dp_phases'))" 2>/dev/null || tail -5 gpurun_out/dp${N}_nccl$mode.log | cut -c1-300
done
