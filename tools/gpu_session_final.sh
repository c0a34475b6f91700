#!/bin/bash
# final-build session A (no profiler): full GPU tests, smoke, default bench (C5 + extra C3), reference arm, the other workloads
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
python -m pytest tests -m gpu -q -p no:cacheprovider -s > gpurun_out/pytest_final_full.log 2>&1
echo "pytest rc=$?" >> gpurun_out/pytest_final_full.log
grep -E "vs oracle|vs float64|FP tensor|boundary sets|KMV |fit vs|passed|failed|FAILED|rc=" gpurun_out/pytest_final_full.log > gpurun_out/pytest_final.log
timeout 300 python __graft_entry__.py smoke > gpurun_out/smoke_final.log 2>&1; echo "smoke rc=$?" >> gpurun_out/smoke_final.log
timeout 900 python bench.py --steps 5 --warmup 3 > gpurun_out/bench_C5_final.json 2> gpurun_out/bench_C5_final.err; echo "bench rc=$?" >> gpurun_out/bench_C5_final.err
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_reference_arm.json 2> /dev/null
for w in C4 C4mf; do
  timeout 600 python bench.py --workload $w --steps 5 --warmup 3 --no-cpu-baseline --no-extra > gpurun_out/bench_${w}_final.json 2> gpurun_out/bench_${w}_final.err
done
for w in C3 C2 C4; do
  timeout 600 python bench.py --workload $w --model parametric --steps 3 --warmup 3 --no-cpu-baseline --no-extra > gpurun_out/bench_${w}_param_final.json 2> /dev/null
done
tail -n 3 gpurun_out/pytest_final.log gpurun_out/smoke_final.log
