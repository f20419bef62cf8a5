#!/bin/bash
mkdir -p gpurun_out
N=2
timeout -s KILL 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29547 bench.py --gpus $N --steps 20 --warmup 5 --skip-cpu --skip-haar --skip-lib > gpurun_out/dp${N}_final.log 2>&1
echo "dp$N exit=$?"
grep -a "unavailable\|Error\|error" gpurun_out/dp${N}_final.log | head -5
grep -a '^{' gpurun_out/dp${N}_final.log | tail -1 | python -c "import json,sys; d=json.loads(sys.stdin.read()); print(round(d['ms_per_step'],3),'ms/step', round(d['value']),'img/s e2e', round(d['e2e']['value']), d['dp_phases'])" || tail -20 gpurun_out/dp${N}_final.log | cut -c1-300
