#!/bin/bash
mkdir -p gpurun_out
run() { name=$1; shift; timeout -s KILL ${T:-300} "$@" > gpurun_out/$name.log 2>&1; echo "$name exit=$?"; tail -n ${TAIL:-6} gpurun_out/$name.log | cut -c1-300; }
T=200 run elementwise python -m pytest tests/test_gpu_elementwise.py -q --tb=short -m gpu -x
T=100 TAIL=10 run gn_micro python tools/gn_microbench.py
T=300 run model python -m pytest tests/test_gpu_model.py -q --tb=short -m gpu
for i in 0 1; do
  timeout 300 python bench.py --steps 30 --warmup 5 --skip-cpu --skip-haar --skip-lib 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('run $i', round(d['ms_per_step'],3), 'ms/step')"
done
timeout 300 python bench.py --config c1 --steps 20 --warmup 5 --skip-lib 2>&1 | tail -1 | cut -c1-1500
