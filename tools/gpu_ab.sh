#!/bin/bash
# A/B of one environment switch on the same box: tools/gpu_ab.sh VAR  (runs VAR=0, VAR=1, VAR=0, VAR=1)
V=$1
for i in 0 1 0 1; do
  env $V=$i python bench.py --steps 30 --warmup 5 --skip-cpu --skip-haar 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('$V=$i', round(d['ms_per_step'],3), 'ms/step')"
done
