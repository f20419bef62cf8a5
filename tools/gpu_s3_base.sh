#!/bin/bash
# session baseline on one GPU: full -m gpu suite, smoke, bench (full line), launch list + dump of conv launch table
mkdir -p gpurun_out
run() { name=$1; shift; timeout -s KILL ${T:-900} "$@" > gpurun_out/$name.log 2>&1; echo "$name exit=$?" | tee -a gpurun_out/summary.txt; tail -n ${TAIL:-3} gpurun_out/$name.log | cut -c1-${CUT:-400}; }
: > gpurun_out/summary.txt
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,memory.total --format=csv > gpurun_out/gpu.txt 2>&1
T=900 run pytest_gpu python -m pytest tests -x -q -m gpu --durations=15
T=200 run smoke python -c "import __graft_entry__ as g; g.smoke()"
T=600 TAIL=1 CUT=8000 run bench python bench.py --dump-kernels
grep -q "bench exit=0" gpurun_out/summary.txt || exit 1
bash tools/gpu_launchlist.sh s3base | head -45
cat gpurun_out/summary.txt
