#!/bin/bash
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
python -m pytest tests -m gpu -q -p no:cacheprovider -s > gpurun_out/pytest2_full.log 2>&1
echo "pytest rc=$?" >> gpurun_out/pytest2_full.log
grep -E "vs oracle|FP tensor|boundary sets|KMV |fit vs|passed|failed|FAILED|rc=" gpurun_out/pytest2_full.log > gpurun_out/pytest2.log
python tools/tensor_errors.py > gpurun_out/tensor_errors.txt 2>&1
timeout 900 python bench.py --steps 5 --warmup 3 > gpurun_out/bench_C5_b.json 2> gpurun_out/bench_C5_b.err
echo "bench rc=$?" >> gpurun_out/bench_C5_b.err
for d in 32 8; do
  PDEIP_LIB=$PWD/pde_inverse_problem_b200/libpdeip_trace.so timeout 300 python tools/tensor_phase_trace.py $d > gpurun_out/trace_d$d.txt 2>&1
done
