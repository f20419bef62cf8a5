#!/bin/bash
# upper bound of what folding the bias / time-row column sums into other kernels can save: step time without them
for d in 0 1 0 1; do
  UB200_DEBUG_SKIP_CHANSUM=$d timeout 300 python bench.py --steps 30 --warmup 5 --skip-cpu --skip-haar --skip-lib 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('SKIP_CHANSUM=$d', round(d['ms_per_step'],3), 'ms/step')"
done
