#!/bin/bash
mkdir -p gpurun_out
timeout 400 python -m pytest tests/test_gpu_model.py -q --tb=short -m gpu -x 2>&1 | tail -3
timeout 300 python bench.py --config c1 --steps 30 --warmup 5 --skip-cpu --skip-haar --skip-lib 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('c1', round(d['ms_per_step'],3), 'ms/step', round(d['value']), d['unit'], 'launches', d.get('gpu_launches'))"
