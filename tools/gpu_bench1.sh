#!/bin/bash
mkdir -p gpurun_out
run() { name=$1; shift; timeout -s KILL ${T:-600} "$@" > gpurun_out/$name.log 2>&1; echo "$name exit=$?" | tee -a gpurun_out/summary.txt; tail -n ${TAIL:-12} gpurun_out/$name.log; }
: > gpurun_out/summary.txt
run elementwise python -m pytest tests/test_gpu_elementwise.py -q --tb=short -m gpu
T=300 TAIL=30 run bench_eager python bench.py --steps 10 --warmup 3 --no-graph --skip-cpu
T=600 TAIL=30 run bench_graph python bench.py --steps 20 --warmup 5
T=300 TAIL=5 run bench_ref python bench.py --impl reference --steps 3 --warmup 1
T=120 run smoke python -c "import __graft_entry__ as g; g.smoke()"
cat gpurun_out/summary.txt
