#!/bin/bash
N=${1:-2}
mkdir -p gpurun_out
run() { name=$1; shift; timeout -s KILL ${T:-600} "$@" > gpurun_out/$name.log 2>&1; echo "$name exit=$?" | tee -a gpurun_out/summary.txt; tail -n ${TAIL:-1} gpurun_out/$name.log | cut -c1-${CUT:-900}; }
: > gpurun_out/summary.txt
T=150 run dp${N}_graph python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29515 bench.py --gpus $N --steps 20 --warmup 5 --skip-cpu --skip-haar
cat gpurun_out/summary.txt
