#!/bin/bash
# final evidence after the wgrad CTA pairs: full -m gpu suite, smoke, bench, launch list, per-launch conv traffic
mkdir -p gpurun_out
run() { name=$1; shift; timeout -s KILL ${T:-900} "$@" > gpurun_out/$name.log 2>&1; echo "$name exit=$?" | tee -a gpurun_out/summary.txt; tail -n ${TAIL:-3} gpurun_out/$name.log | cut -c1-${CUT:-400}; }
: > gpurun_out/summary.txt
T=900 run pytest_gpu python -m pytest tests -x -q -m gpu
T=200 run smoke python -c "import __graft_entry__ as g; g.smoke()"
T=600 TAIL=1 CUT=9000 run bench python bench.py --dump-kernels
grep -q "bench exit=0" gpurun_out/summary.txt || exit 1
for c in c3 c4; do T=300 TAIL=1 CUT=200 run bench_$c python bench.py --config $c --skip-cpu --skip-lib; done
bash tools/gpu_launchlist.sh final2 | head -16
bash tools/gpu_conv_traffic.sh
cat gpurun_out/summary.txt
