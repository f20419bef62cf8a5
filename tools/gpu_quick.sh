#!/bin/bash
# quick loop: elementwise + conv + model parity, then one graph-mode bench line
mkdir -p gpurun_out
run() { name=$1; shift; timeout -s KILL ${T:-600} "$@" > gpurun_out/$name.log 2>&1; echo "$name exit=$?" | tee -a gpurun_out/summary.txt; tail -n ${TAIL:-6} gpurun_out/$name.log; }
: > gpurun_out/summary.txt
run elementwise python -m pytest tests/test_gpu_elementwise.py -q --tb=short -m gpu
T=420 run conv python -m pytest tests/test_gpu_conv.py -q --tb=short -m gpu
T=420 run model python -m pytest tests/test_gpu_model.py -q --tb=short -m gpu
T=600 TAIL=3 run bench python bench.py --steps 20 --warmup 5 --skip-cpu --skip-haar
cat gpurun_out/summary.txt
