#!/bin/bash
# conv-kernel iteration loop: conv + model parity, then bench with the per-launch conv table
mkdir -p gpurun_out
run() { name=$1; shift; timeout -s KILL ${T:-600} "$@" > gpurun_out/$name.log 2>&1; echo "$name exit=$?" | tee -a gpurun_out/summary.txt; tail -n ${TAIL:-4} gpurun_out/$name.log | cut -c1-600; }
: > gpurun_out/summary.txt
T=300 run conv python -m pytest tests/test_gpu_conv.py -q --tb=short -m gpu -x
T=300 run model python -m pytest tests/test_gpu_model.py -q --tb=short -m gpu -x
T=300 TAIL=1 run bench python bench.py --steps 20 --warmup 5 --skip-cpu --skip-haar --dump-kernels
cat gpurun_out/summary.txt
