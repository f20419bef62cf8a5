#!/bin/bash
# oversized GroupNorm slabs through the streaming pair (c3 / c4 / c5 A/B), GN parity tests, then ncu --set full captures of the
# wgrad kernel with the TMA reduce-add epilogue and of the streaming GroupNorm backward kernels inside one eager config-2 step
mkdir -p gpurun_out
timeout 400 python -m pytest tests/test_gpu_elementwise.py tests/test_gpu_model.py -q --tb=short -m gpu -x 2>&1 | tail -3
for c in c3 c4 c5; do
  for s in 0 x; do
    if [ $s = x ]; then unset UB200_GN_STREAM; else export UB200_GN_STREAM=$s; fi
    timeout 300 python bench.py --config $c --steps 30 --warmup 5 --skip-cpu --skip-haar --skip-lib 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('$c GN_STREAM=$s', round(d['ms_per_step'],3), 'ms/step', round(d['value']), d['unit'])"
  done
done
unset UB200_GN_STREAM
CMD="python bench.py --steps 1 --warmup 3 --no-graph --skip-cpu --skip-haar --skip-lib --profile-step"
timeout -s KILL 300 $CMD > gpurun_out/plain.log 2>&1 || { echo plain failed; tail -5 gpurun_out/plain.log; exit 1; }
timeout -s KILL 600 ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:conv_wgrad_kernel -c 8 \
   -o gpurun_out/r02_prof_wgrad_tmared -f $CMD > gpurun_out/ncu_wgrad.log 2>&1; echo "ncu wgrad exit=$?"
timeout -s KILL 600 ncu --set full --clock-control none --profile-from-start off -k regex:gn_stream_bwd -c 8 \
   -o gpurun_out/r02_prof_gn_stream_bwd -f $CMD > gpurun_out/ncu_gn.log 2>&1; echo "ncu gn exit=$?"
ls -la gpurun_out/*.ncu-rep
