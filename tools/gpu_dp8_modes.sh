#!/bin/bash
mkdir -p gpurun_out
N=${N:-8}
b() { name=$1; timeout -s KILL 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $((29500 + RANDOM % 400)) bench.py --gpus $N --steps 30 --warmup 5 --skip-cpu --skip-haar --skip-lib > gpurun_out/dp${N}_$name.log 2>&1
  echo -n "$name exit=$? "; grep -a '^{' gpurun_out/dp${N}_$name.log | tail -1 | python -c "import json,sys; d=json.loads(sys.stdin.read()); print(round(d['ms_per_step'],3),'ms/step', round(d['value']),'img/s', d.get('dp_phases'))" 2>/dev/null || tail -8 gpurun_out/dp${N}_$name.log | cut -c1-300; }
UB200_DP_P2P=0 b nccl_tail
UB200_DP_OVERLAP=0 b p2p_tail
UB200_P2P_EARLY_CTAS=32 b p2p_overlap_early32
