#!/bin/bash
# conv_wgrad CTA pairs on the non-halo paths too (8x8 3x3 layers, 1x1 convs): UB200_WGRAD_PAIR=2 vs 1
mkdir -p gpurun_out
L=gpurun_out/wgrad_pair2.log; : > $L
sed -i 's/^shapes = \[/shapes = [(4, 8, 8, 256, 256, 3), (2, 16, 16, 256, 512, 1), (3, 8, 8, 384, 256, 1), /' tools/wgrad_probe.py
for d in 1 2; do
  UB200_WGRAD_PAIR=$d timeout 150 python tools/wgrad_probe.py 2>&1 | grep -E "time|check|worst|rror" | sed "s/^/PAIR=$d /" >> $L
  echo "probe PAIR=$d exit=${PIPESTATUS[0]}" >> $L
done
grep -E "FAIL|worst|rror|exit|8x8|k1" $L
if grep -q "PAIR=2 .*worst rel err [0-9.]*e-0[5-9]" $L; then
  UB200_WGRAD_PAIR=2 timeout 300 python -m pytest tests/test_gpu_conv.py -q --tb=short -m gpu -x 2>&1 | tail -3
  for d in 1 2 1 2; do
    UB200_WGRAD_PAIR=$d timeout 300 python bench.py --steps 30 --warmup 5 --skip-cpu --skip-haar --skip-lib 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('WGRAD_PAIR=$d', round(d['ms_per_step'],3), 'ms/step')"
  done
fi
