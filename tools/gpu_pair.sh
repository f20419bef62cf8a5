#!/bin/bash
# cta_group::2 pair mode of the conv kernel: parity under a hard timeout (a protocol bug would hang), microbench and step A/B
mkdir -p gpurun_out
run() { name=$1; shift; timeout -s KILL ${T:-300} "$@" > gpurun_out/$name.log 2>&1; echo "$name exit=$?"; tail -n ${TAIL:-6} gpurun_out/$name.log | cut -c1-300; }
UB200_FPROP_PAIR=1 T=240 run conv_pair python -m pytest tests/test_gpu_conv.py -q --tb=short -m gpu -x
grep -q "passed" gpurun_out/conv_pair.log && ! grep -q "failed" gpurun_out/conv_pair.log || { echo "pair mode parity failed"; exit 0; }
UB200_FPROP_PAIR=1 T=300 run model_pair python -m pytest tests/test_gpu_model.py -q --tb=short -m gpu -x
UB200_FPROP_PAIR=0 T=120 TAIL=16 run micro_pair0 python tools/conv_microbench.py
UB200_FPROP_PAIR=1 T=120 TAIL=16 run micro_pair1 python tools/conv_microbench.py
for i in 0 1 0 1; do
  UB200_FPROP_PAIR=$i timeout 300 python bench.py --steps 30 --warmup 5 --skip-cpu --skip-haar --skip-lib 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('PAIR=$i', round(d['ms_per_step'],3), 'ms/step', 'fprop', round(d['roofline']['achieved'],1))"
done
