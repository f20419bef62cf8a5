#!/bin/bash
# in-graph A/B of the streaming-GroupNorm policy (per-sample slab window, direction)
b() { timeout 300 python bench.py --steps 30 --warmup 5 --skip-cpu --skip-haar --skip-lib 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('$1', round(d['ms_per_step'],3), 'ms/step')"; }
b all_a
b all_b
UB200_GN_STREAM_MAX_SLAB=600000 b max600k
UB200_GN_STREAM_MAX_SLAB=400000 b max400k
UB200_GN_STREAM_MIN_SLAB=40000 b min40k
UB200_GN_STREAM_DIR=1 b fwd_only
UB200_GN_STREAM_DIR=2 b bwd_only
UB200_GN_STREAM=0 b none
