#!/bin/bash
# conv_wgrad CTA pairs (cta_group::2): correctness on the probe shapes, timing, conv tests, step A/B
mkdir -p gpurun_out
L=gpurun_out/wgrad_pair.log; : > $L
for d in 0 1; do
  UB200_WGRAD_PAIR=$d timeout 150 python tools/wgrad_probe.py 2>&1 | grep -E "time|check|worst|rror" | sed "s/^/PAIR=$d /" >> $L
  echo "probe PAIR=$d exit=${PIPESTATUS[0]}" >> $L
done
grep -E "FAIL|worst|time|rror|exit" $L
if grep -q "PAIR=1 .*worst rel err [0-9.]*e-0[5-9]" $L; then
  UB200_WGRAD_PAIR=1 timeout 300 python -m pytest tests/test_gpu_conv.py -q --tb=short -m gpu -x 2>&1 | tail -3
  for d in 0 1 0 1; do
    UB200_WGRAD_PAIR=$d timeout 300 python bench.py --steps 30 --warmup 5 --skip-cpu --skip-haar --skip-lib 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('WGRAD_PAIR=$d', round(d['ms_per_step'],3), 'ms/step')"
  done
fi
