#!/bin/bash
mkdir -p gpurun_out
run() { name=$1; shift; timeout -s KILL ${T:-600} "$@" > gpurun_out/$name.log 2>&1; echo "$name exit=$?"; tail -n ${TAIL:-3} gpurun_out/$name.log | cut -c1-${CUT:-1500}; }
T=900 TAIL=12 run pytest_gpu python -m pytest tests -q --tb=short -m gpu -p no:cacheprovider
T=400 TAIL=1 CUT=9000 run bench_c2 python bench.py --steps 20 --warmup 5
for c in c1 c3 c3u c4 c5; do T=400 TAIL=1 CUT=3000 run bench_$c python bench.py --config $c --steps 20 --warmup 5; done
T=600 TAIL=25 run loss_curve0 python tools/loss_curve.py --steps 1000 --batch 128 --dropout 0.0 --out gpurun_out/r02_loss_curve_dropout0.json
T=600 TAIL=25 run loss_curve1 python tools/loss_curve.py --steps 1000 --batch 128 --dropout 0.1 --out gpurun_out/r02_loss_curve_dropout01.json
