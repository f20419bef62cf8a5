#!/bin/bash
timeout 300 python -m pytest tests/test_gpu_conv.py tests/test_gpu_model.py -q --tb=short -m gpu -x 2>&1 | tail -4
for c in c2 c3; do
timeout 300 python bench.py --config $c --steps 30 --warmup 5 --skip-cpu --skip-haar --skip-lib 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('$c', round(d['ms_per_step'],3), 'ms/step', round(d['value']))"
done
