#!/bin/bash
# data-parallel bench on N GPUs of one box: eager first (safe), then whole-step CUDA graph
N=${1:-2}
mkdir -p gpurun_out
run() { name=$1; shift; timeout -s KILL ${T:-600} "$@" > gpurun_out/$name.log 2>&1; echo "$name exit=$?" | tee -a gpurun_out/summary.txt; tail -n ${TAIL:-4} gpurun_out/$name.log | cut -c1-1500; }
: > gpurun_out/summary.txt
T=150 run dp${N}_eager python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 10 --warmup 3 --no-graph --skip-cpu --skip-haar
T=150 run dp${N}_graph python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus $N --steps 20 --warmup 5 --skip-cpu --skip-haar
cat gpurun_out/summary.txt
