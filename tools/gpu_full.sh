#!/bin/bash
# everything the driver runs at round end: the whole -m gpu suite, smoke(), both bench arms
mkdir -p gpurun_out
run() { name=$1; shift; timeout -s KILL ${T:-900} "$@" > gpurun_out/$name.log 2>&1; echo "$name exit=$?" | tee -a gpurun_out/summary.txt; tail -n ${TAIL:-3} gpurun_out/$name.log | cut -c1-${CUT:-400}; }
: > gpurun_out/summary.txt
T=900 run pytest_gpu python -m pytest tests -x -q -m gpu
T=200 run smoke python -c "import __graft_entry__ as g; g.smoke()"
T=600 TAIL=1 CUT=6000 run bench python bench.py
T=300 TAIL=1 CUT=1200 run bench_ref python bench.py --impl reference --steps 3 --warmup 1
cat gpurun_out/summary.txt
