#!/bin/bash
mkdir -p gpurun_out
run() { name=$1; shift; timeout -s KILL ${T:-300} "$@" > gpurun_out/$name.log 2>&1; echo "$name exit=$?"; tail -n ${TAIL:-6} gpurun_out/$name.log | cut -c1-300; }
T=900 TAIL=6 run pytest_gpu python -m pytest tests -q --tb=short -m gpu -p no:cacheprovider
bash tools/gpu_launchlist.sh r2b | head -40
