#!/bin/bash
# final-build session: full GPU tests, smoke, default bench, ncu launch lists + full captures (C5 and C3 residual kernels)
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
python -m pytest tests -m gpu -q -p no:cacheprovider -s > gpurun_out/pytest7_full.log 2>&1
echo "pytest rc=$?" >> gpurun_out/pytest7_full.log
grep -E "vs oracle|vs float64|FP tensor|boundary sets|KMV |fit vs|passed|failed|FAILED|rc=" gpurun_out/pytest7_full.log > gpurun_out/pytest7.log
timeout 300 python __graft_entry__.py smoke > gpurun_out/smoke7.log 2>&1; echo "smoke rc=$?" >> gpurun_out/smoke7.log
timeout 900 python bench.py --steps 5 --warmup 3 > gpurun_out/bench_C5_final.json 2> gpurun_out/bench_C5_final.err; echo "bench rc=$?" >> gpurun_out/bench_C5_final.err
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_reference_arm.json 2> /dev/null
CMD5="python bench.py --workload C5 --particles 113664 --steps 1 --warmup 3 --no-cpu-baseline --no-extra"
CMD3="python bench.py --workload C3 --particles 606208 --steps 1 --warmup 3 --no-cpu-baseline"
$CMD5 > gpurun_out/plain5.log 2>&1 && $CMD3 > gpurun_out/plain3.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/launches_C5_final.csv $CMD5 > gpurun_out/ncu_a.log 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/launches_C3_final.csv $CMD3 > gpurun_out/ncu_b.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:mlp_residual_tc -s 6 -c 1 -o gpurun_out/prof_res8 $CMD3 > gpurun_out/ncu_c.log 2>&1
ls -la gpurun_out | tail -20
