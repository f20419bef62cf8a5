#!/bin/bash
mkdir -p gpurun_out
run() { name=$1; shift; timeout -s KILL ${T:-300} "$@" > gpurun_out/$name.log 2>&1; echo "$name exit=$?"; tail -n ${TAIL:-6} gpurun_out/$name.log | cut -c1-300; }
T=120 TAIL=25 run attn python -m pytest tests/test_gpu_attention.py -q --tb=short -m gpu -x
grep -q "passed" gpurun_out/attn.log && ! grep -q "failed" gpurun_out/attn.log || exit 0
T=400 TAIL=8 run model python -m pytest tests/test_gpu_model.py -q --tb=short -m gpu
for i in 0 1 0 1; do
  UB200_ATTN_CORE=$i timeout 300 python bench.py --steps 30 --warmup 5 --skip-cpu --skip-haar --skip-lib 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('ATTN_CORE=$i', round(d['ms_per_step'],3), 'ms/step')"
done
