#!/bin/bash
for p in 0 2 0 2; do UB200_WGRAD_PAIR=$p timeout 60 python tools/wgrad_power_probe.py 2>&1 | tail -2; done
