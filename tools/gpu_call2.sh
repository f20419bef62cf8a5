#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -q --tb=short -m gpu -p no:cacheprovider > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit=$?"; tail -n 30 gpurun_out/pytest_gpu.log
cat gpurun_out/config2_parity.json
bash tools/gpu_launchlist.sh r2a | head -45
