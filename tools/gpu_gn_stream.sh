#!/bin/bash
# streaming two-launch GroupNorm vs the cluster kernels: parity with the stream forced on, microbench both, step A/B
mkdir -p gpurun_out
run() { name=$1; shift; timeout -s KILL ${T:-300} "$@" > gpurun_out/$name.log 2>&1; echo "$name exit=$?"; tail -n ${TAIL:-6} gpurun_out/$name.log | cut -c1-300; }
UB200_GN_STREAM=1 T=300 run elementwise_stream1 python -m pytest tests/test_gpu_elementwise.py -q --tb=short -m gpu -x
T=300 run elementwise_policy python -m pytest tests/test_gpu_elementwise.py -q --tb=short -m gpu -x
UB200_GN_STREAM=0 T=100 TAIL=10 run gn_micro_stream0 python tools/gn_microbench.py
UB200_GN_STREAM=1 T=100 TAIL=10 run gn_micro_stream1 python tools/gn_microbench.py
T=300 run model python -m pytest tests/test_gpu_model.py -q --tb=short -m gpu -x
for s in 0 1 p; do
  if [ $s = p ]; then unset UB200_GN_STREAM; else export UB200_GN_STREAM=$s; fi
  timeout 300 python bench.py --steps 30 --warmup 5 --skip-cpu --skip-haar --skip-lib 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('stream=$s', round(d['ms_per_step'],3), 'ms/step')"
done
