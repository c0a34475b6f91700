#!/bin/bash
# one profiler step per call (the same command first runs plain and must exit 0):
#   tools/gpu_session_ncu.sh launches C5 | launches C3 | full C5 | full C3
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
kind=$1; w=$2
if [ "$w" = C5 ]; then CMD="python bench.py --workload C5 --particles 113664 --steps 1 --warmup 3 --no-cpu-baseline --no-extra"
else CMD="python bench.py --workload C3 --particles 606208 --steps 1 --warmup 3 --no-cpu-baseline"; fi
$CMD > gpurun_out/plain_$w.log 2>&1 || { echo "plain run failed"; exit 1; }
if [ "$kind" = launches ]; then
  ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/launches_${w}_final.csv $CMD > gpurun_out/ncu_l_$w.log 2>&1
else
  ncu --set full --clock-control none --import-source on -k regex:mlp_residual_tc -s 4 -c 1 -f -o gpurun_out/prof_res_${w}_final $CMD > gpurun_out/ncu_f_$w.log 2>&1
fi
ls -la gpurun_out | tail -5
