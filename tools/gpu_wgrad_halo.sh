#!/bin/bash
# round-2 call 1: wgrad halo-box probe (correctness of each descriptor base-offset mode, then timing + step time of the mode
# that passes), then the whole -m gpu suite and one bench line
mkdir -p gpurun_out
L=gpurun_out/wgrad_halo.log; : > $L
UB200_WGRAD_HALO=0 timeout 300 python tools/wgrad_probe.py >> $L 2>&1
PASS=""
for bo in 0 1 2; do
  PROBE_TIME=0 UB200_WGRAD_HALO=1 UB200_WGRAD_BO=$bo timeout 120 python tools/wgrad_probe.py > gpurun_out/halo_bo$bo.log 2>&1
  cat gpurun_out/halo_bo$bo.log >> $L
  if grep -q "worst rel err" gpurun_out/halo_bo$bo.log && ! grep -q FAIL gpurun_out/halo_bo$bo.log; then
    echo "MODE bo=$bo PASSES" >> $L
    [ -z "$PASS" ] && PASS=$bo
  fi
done
if [ -n "$PASS" ]; then
  UB200_WGRAD_HALO=1 UB200_WGRAD_BO=$PASS timeout 300 python tools/wgrad_probe.py >> $L 2>&1
  for h in 0 1 0 1; do
    UB200_WGRAD_HALO=$h UB200_WGRAD_BO=$PASS timeout 300 python bench.py --steps 30 --warmup 5 --skip-cpu --skip-haar 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('HALO=$h BO=$PASS', round(d['ms_per_step'],3), 'ms/step')" >> $L 2>&1
  done
  export UB200_WGRAD_HALO=1 UB200_WGRAD_BO=$PASS
else
  export UB200_WGRAD_HALO=0
fi
grep -E "worst|PASSES|ms/step|time" $L | tail -70
echo "== gpu tests (UB200_WGRAD_HALO=$UB200_WGRAD_HALO UB200_WGRAD_BO=$UB200_WGRAD_BO)"
timeout 1500 python -m pytest tests -q --tb=short -m gpu -x -p no:cacheprovider > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit=$?"; tail -n 25 gpurun_out/pytest_gpu.log
