#!/bin/bash
# round-2 final evidence on one GPU: full -m gpu suite, smoke, both bench arms, the other configs, the one-step launch list
mkdir -p gpurun_out
run() { name=$1; shift; timeout -s KILL ${T:-900} "$@" > gpurun_out/$name.log 2>&1; echo "$name exit=$?" | tee -a gpurun_out/summary.txt; tail -n ${TAIL:-3} gpurun_out/$name.log | cut -c1-${CUT:-400}; }
: > gpurun_out/summary.txt
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,memory.total --format=csv > gpurun_out/gpu.txt 2>&1
T=900 run pytest_gpu python -m pytest tests -x -q -m gpu -rs
T=200 run smoke python -c "import __graft_entry__ as g; g.smoke()"
T=600 TAIL=1 CUT=9000 run bench python bench.py --dump-kernels
T=400 TAIL=1 CUT=1500 run bench_ref python bench.py --impl reference --steps 3 --warmup 1
grep -q "bench exit=0" gpurun_out/summary.txt || exit 1
for c in c1 c3 c3u c4 c5; do
  T=400 TAIL=1 CUT=300 run bench_$c python bench.py --config $c
done
bash tools/gpu_launchlist.sh final | head -40
cat gpurun_out/summary.txt
