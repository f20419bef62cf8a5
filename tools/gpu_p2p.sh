#!/bin/bash
# N-GPU: peer-memory all-reduce checks, then the data-parallel bench with it (default) and with NCCL after the graph
mkdir -p gpurun_out
N=${N:-2}
tr() { timeout -s KILL ${T:-300} python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $((29500 + RANDOM % 500)) "$@"; }
tr tools/p2p_check.py > gpurun_out/p2p_check_n$N.log 2>&1; echo "p2p_check exit=$?"; grep -a '^{' gpurun_out/p2p_check_n$N.log | tail -1 | cut -c1-1500 || tail -20 gpurun_out/p2p_check_n$N.log
grep -a -i "error\|Traceback\|unavailable" gpurun_out/p2p_check_n$N.log | head -10
for mode in 1 0; do
  UB200_DP_P2P=$mode T=400 tr bench.py --gpus $N --steps 30 --warmup 5 --skip-cpu --skip-haar --skip-lib > gpurun_out/dp${N}_p2p$mode.log 2>&1
  echo "DP_P2P=$mode exit=$?"
  grep -a "unavailable" gpurun_out/dp${N}_p2p$mode.log | head -3
  grep -a '^{' gpurun_out/dp${N}_p2p$mode.log | tail -1 | python -c "import json,sys; d=json.loads(sys.stdin.read()); print(round(d['ms_per_step'],3),'ms/step', round(d['value']),'img/s', d.get('dp_phases'), d['config'].get('allreduce'))" 2>/dev/null || tail -8 gpurun_out/dp${N}_p2p$mode.log | cut -c1-300
done
