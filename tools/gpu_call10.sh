#!/bin/bash
mkdir -p gpurun_out
run() { name=$1; shift; timeout -s KILL ${T:-300} "$@" > gpurun_out/$name.log 2>&1; echo "$name exit=$?"; tail -n ${TAIL:-6} gpurun_out/$name.log | cut -c1-300; }
T=900 TAIL=8 run pytest_gpu python -m pytest tests -q --tb=short -m gpu -p no:cacheprovider
UB200_GN_BWD_PACKED=0 T=100 TAIL=6 run gn_packed0 python tools/gn_microbench.py
UB200_GN_BWD_PACKED=1 T=100 TAIL=6 run gn_packed1 python tools/gn_microbench.py
for i in 0 1 0 1; do
  UB200_GN_BWD_PACKED=$i timeout 300 python bench.py --steps 30 --warmup 5 --skip-cpu --skip-haar --skip-lib 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('GN_BWD_PACKED=$i', round(d['ms_per_step'],3), 'ms/step', 'fprop', round(d['roofline']['achieved'],1), 'wgrad', round(d['roofline']['wgrad']['achieved'],1))"
done
