#!/bin/bash
# N-GPU data-parallel bench with NCCL's algorithm choice logged
N=${1:-8}
mkdir -p gpurun_out
run() { name=$1; shift; timeout -s KILL ${T:-600} "$@" > gpurun_out/$name.log 2>&1; echo "$name exit=$?" | tee -a gpurun_out/summary.txt; tail -n ${TAIL:-1} gpurun_out/$name.log | cut -c1-${CUT:-1500}; }
: > gpurun_out/summary.txt
NCCL_DEBUG=INFO NCCL_DEBUG_SUBSYS=INIT,COLL T=200 run dp${N}_graph python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29515 bench.py --gpus $N --steps 20 --warmup 5 --skip-cpu --skip-haar
grep -m5 -E "NVLS|Algo|algorithm|Ring|Tree" gpurun_out/dp${N}_graph.log | cut -c1-300
grep -c "AllReduce" gpurun_out/dp${N}_graph.log
grep -m3 "AllReduce" gpurun_out/dp${N}_graph.log | cut -c1-400
cat gpurun_out/summary.txt
