#!/bin/bash
mkdir -p gpurun_out
run() { name=$1; shift; timeout -s KILL ${T:-300} "$@" > gpurun_out/$name.log 2>&1; echo "$name exit=$?" | tee -a gpurun_out/summary.txt; tail -n ${TAIL:-6} gpurun_out/$name.log; }
: > gpurun_out/summary.txt
T=300 run conv python -m pytest tests/test_gpu_conv.py -q --tb=short -m gpu -x
grep -q "conv exit=0" gpurun_out/summary.txt || exit 1
T=120 TAIL=16 run micro_fast python tools/conv_microbench.py
T=300 run model python -m pytest tests/test_gpu_model.py -q --tb=short -m gpu
T=300 TAIL=1 run bench python bench.py --steps 20 --warmup 5 --skip-cpu --skip-haar
cat gpurun_out/summary.txt
