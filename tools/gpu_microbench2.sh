#!/bin/bash
echo "== msub=2"; UB200_FPROP_MSUB=2 timeout 120 python tools/conv_microbench.py 2>&1 | grep ^BN
echo "== cluster=2"; UB200_FPROP_CLUSTER=2 timeout 120 python tools/conv_microbench.py 2>&1 | grep ^BN
echo "== cluster=2 msub=2"; UB200_FPROP_CLUSTER=2 UB200_FPROP_MSUB=2 timeout 120 python tools/conv_microbench.py 2>&1 | grep ^BN
echo "== stages 3"; UB200_FPROP_STAGES=3 timeout 120 python tools/conv_microbench.py 2>&1 | grep ^BN
