#!/bin/bash
# Run on the GPU box (via gpurun): each suite in its own process under a hard timeout so that a hung kernel
# cannot take the whole call down.  Logs land in gpurun_out/.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,memory.total --format=csv > gpurun_out/gpu.txt 2>&1
run() { name=$1; shift; timeout -s KILL ${T:-600} "$@" > gpurun_out/$name.log 2>&1; echo "$name exit=$?" | tee -a gpurun_out/summary.txt; tail -n ${TAIL:-15} gpurun_out/$name.log; }
: > gpurun_out/summary.txt
run haar python -m pytest tests/test_gpu_haar.py -q --tb=short -m gpu
run elementwise python -m pytest tests/test_gpu_elementwise.py -q --tb=short -m gpu
T=420 run conv python -m pytest tests/test_gpu_conv.py -q --tb=short -m gpu
T=420 run model python -m pytest tests/test_gpu_model.py -q --tb=short -m gpu
cat gpurun_out/summary.txt
