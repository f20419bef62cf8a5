#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests -x -q -m gpu 2>&1 | tail -2
timeout 200 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
timeout 400 python bench.py > gpurun_out/bench_last.log 2>&1; echo "bench exit=$?"
grep -a '^{' gpurun_out/bench_last.log | tail -1 | python -c "import json,sys; d=json.loads(sys.stdin.read()); print(round(d['ms_per_step'],3),'ms/step', round(d['value']),'img/s e2e', round(d['e2e']['value']), 'launches', d['gpu_launches'], d['clocks'])"
