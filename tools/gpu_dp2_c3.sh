#!/bin/bash
mkdir -p gpurun_out
timeout -s KILL 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29551 bench.py --gpus 2 --config c3 --steps 20 --warmup 5 --skip-cpu --skip-lib > gpurun_out/dp2_c3.log 2>&1
echo "dp2 c3 exit=$?"
grep -a '^{' gpurun_out/dp2_c3.log | tail -1 | python -c "import json,sys; d=json.loads(sys.stdin.read()); print(round(d['ms_per_step'],3),'ms/step', round(d['value']), d['unit'], 'e2e', round(d['e2e']['value']))" || tail -15 gpurun_out/dp2_c3.log | cut -c1-300
