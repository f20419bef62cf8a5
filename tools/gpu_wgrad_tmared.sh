#!/bin/bash
# TMA reduce-add epilogue of conv_wgrad vs the per-thread atomics: correctness on the probe shapes, timing, conv tests, step A/B
mkdir -p gpurun_out
L=gpurun_out/wgrad_tmared.log; : > $L
for d in 0 1; do
  UB200_WGRAD_TMA_RED=$d timeout 200 python tools/wgrad_probe.py 2>&1 | grep -E "time|check|worst|rror" | sed "s/^/TMA_RED=$d /" >> $L
done
grep -E "FAIL|worst|time|rror" $L
timeout 400 python -m pytest tests/test_gpu_conv.py tests/test_gpu_model.py -q --tb=short -m gpu -x 2>&1 | tail -5
for d in 0 1 0 1; do
  UB200_WGRAD_TMA_RED=$d timeout 300 python bench.py --steps 30 --warmup 5 --skip-cpu --skip-haar --skip-lib 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('TMA_RED=$d', round(d['ms_per_step'],3), 'ms/step')"
done
