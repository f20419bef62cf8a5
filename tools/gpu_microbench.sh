#!/bin/bash
for st in 2 4 8; do UB200_FPROP_STAGES=$st timeout 120 python tools/conv_microbench.py; done
for bn in 64 128 256; do UB200_FPROP_BN=$bn timeout 120 python tools/conv_microbench.py; done
