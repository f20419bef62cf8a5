#!/bin/bash
# round-end evidence on one GPU: full -m gpu suite, smoke, both bench arms, then (only after those exited 0 without
# ncu) the one-step launch list and one `ncu --set full` capture of the largest conv_fprop launches
mkdir -p gpurun_out
run() { name=$1; shift; timeout -s KILL ${T:-900} "$@" > gpurun_out/$name.log 2>&1; echo "$name exit=$?" | tee -a gpurun_out/summary.txt; tail -n ${TAIL:-3} gpurun_out/$name.log | cut -c1-${CUT:-400}; }
: > gpurun_out/summary.txt
T=900 run pytest_gpu python -m pytest tests -x -q -m gpu
T=200 run smoke python -c "import __graft_entry__ as g; g.smoke()"
T=600 TAIL=1 CUT=6000 run bench python bench.py
T=300 TAIL=1 CUT=1200 run bench_ref python bench.py --impl reference --steps 3 --warmup 1
grep -q "bench exit=0" gpurun_out/summary.txt || exit 1
bash tools/gpu_launchlist.sh final | head -30
timeout -s KILL 600 ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:conv_fprop_kernel -s 32 -c 8 \
  -o gpurun_out/prof_conv_fprop_final -f python bench.py --steps 1 --warmup 3 --no-graph --skip-cpu --skip-haar --profile-step > gpurun_out/ncu_full_conv.log 2>&1
echo ncu_full exit=$?
cat gpurun_out/summary.txt
