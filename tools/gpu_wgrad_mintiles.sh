#!/bin/bash
# with the cheaper TMA reduce-add epilogue, fewer 64-pixel stages per split may pay: sweep UB200_WGRAD_MIN_TILES
mkdir -p gpurun_out
for m in 6 4 3 2; do
  UB200_WGRAD_MIN_TILES=$m timeout 300 python bench.py --steps 30 --warmup 5 --skip-cpu --skip-haar --skip-lib 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('MIN_TILES=$m', round(d['ms_per_step'],3), 'ms/step')"
done
for m in 6 3 2; do
  UB200_WGRAD_MIN_TILES=$m timeout 200 python tools/wgrad_probe.py 2>&1 | grep -E "time +(4x4|8x8|16x16)" | sed "s/^/MIN_TILES=$m /"
done
