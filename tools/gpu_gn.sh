#!/bin/bash
# GroupNorm iteration loop: parity of the elementwise kernels, then the GN microbench (fused vs multi-pass), then one bench line
mkdir -p gpurun_out
run() { name=$1; shift; timeout -s KILL ${T:-300} "$@" > gpurun_out/$name.log 2>&1; echo "$name exit=$?" | tee -a gpurun_out/summary.txt; tail -n ${TAIL:-6} gpurun_out/$name.log; }
: > gpurun_out/summary.txt
T=150 run elementwise python -m pytest tests/test_gpu_elementwise.py -q --tb=short -m gpu -x
grep -q "exit=0" gpurun_out/summary.txt || exit 1
T=100 TAIL=12 run gn_fused python tools/gn_microbench.py
T=300 run model python -m pytest tests/test_gpu_model.py -q --tb=short -m gpu
T=300 TAIL=1 run bench python bench.py --steps 20 --warmup 5 --skip-cpu --skip-haar
cat gpurun_out/summary.txt
