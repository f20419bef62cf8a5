#!/bin/bash
mkdir -p gpurun_out
N=${N:-2}
tr() { timeout -s KILL ${T:-300} python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $((29500 + RANDOM % 500)) "$@"; }
b() { name=$1; T=400 tr bench.py --gpus $N --steps 30 --warmup 5 --skip-cpu --skip-haar --skip-lib > gpurun_out/dp${N}_$name.log 2>&1
  echo -n "$name exit=$? "; grep -a '^{' gpurun_out/dp${N}_$name.log | tail -1 | python -c "import json,sys; d=json.loads(sys.stdin.read()); print(round(d['ms_per_step'],3),'ms/step', round(d['value']),'img/s', d.get('dp_phases'))" 2>/dev/null || tail -8 gpurun_out/dp${N}_$name.log | cut -c1-300; }
tr tools/p2p_check.py > gpurun_out/p2p_check_n$N.log 2>&1; echo "p2p_check exit=$?"; grep -a '^{' gpurun_out/p2p_check_n$N.log | tail -1 | python -c "import json,sys; d=json.loads(sys.stdin.read()); print({k:v for k,v in d.items() if not k.startswith('trial')})"
UB200_P2P_EARLY_CTAS=16 b early16
UB200_P2P_EARLY_CTAS=32 b early32
UB200_P2P_EARLY_CTAS=64 b early64
UB200_P2P_EVICT_FIRST=1 b ef_early148
UB200_P2P_EVICT_FIRST=1 UB200_P2P_EARLY_CTAS=32 b ef_early32
UB200_P2P_EARLY_CTAS=32 UB200_DP_BUCKET_MB=32 b early32_bucket32
UB200_DP_OVERLAP=0 b tail
