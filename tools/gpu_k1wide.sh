#!/bin/bash
mkdir -p gpurun_out
run() { name=$1; shift; timeout -s KILL ${T:-300} "$@" > gpurun_out/$name.log 2>&1; echo "$name exit=$?"; tail -n ${TAIL:-6} gpurun_out/$name.log | cut -c1-300; }
T=200 TAIL=4 run conv python -m pytest tests/test_gpu_conv.py -q --tb=short -m gpu -x
UB200_WGRAD_K1WIDE=0 T=120 TAIL=3 run probe0 python tools/wgrad_probe.py
UB200_WGRAD_K1WIDE=1 T=120 TAIL=3 run probe1 python tools/wgrad_probe.py
for i in 0 1 0 1; do
  UB200_WGRAD_K1WIDE=$i timeout 300 python bench.py --steps 30 --warmup 5 --skip-cpu --skip-haar --skip-lib 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('K1WIDE=$i', round(d['ms_per_step'],3), 'ms/step', 'wgrad', round(d['roofline']['wgrad']['achieved'],1))"
done
