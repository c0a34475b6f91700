#!/bin/bash
# round-2 GPU session 1: full GPU test suite, lo-halves accuracy study, default bench (C5 + extra.c3), smoke, ncu
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv > gpurun_out/smi.txt 2>&1
python -m pytest tests -m gpu -q -p no:cacheprovider -s 2>&1 | tail -150 > gpurun_out/pytest1.log
echo "pytest rc=$?" >> gpurun_out/pytest1.log
for v in "" _nlo32 _nlo12 _nlo31; do
  echo "=== libpdeip$v.so" >> gpurun_out/nlo.log
  PDEIP_LIB=$PWD/pde_inverse_problem_b200/libpdeip$v.so timeout 300 python tools/tensor_errors.py >> gpurun_out/nlo.log 2>&1
done
timeout 900 python bench.py --steps 5 --warmup 3 > gpurun_out/bench_C5.json 2> gpurun_out/bench_C5.err
echo "bench rc=$?" >> gpurun_out/bench_C5.err
timeout 300 python __graft_entry__.py smoke > gpurun_out/smoke.log 2>&1
echo "smoke rc=$?" >> gpurun_out/smoke.log
CMD="python bench.py --workload C5 --particles 113664 --steps 1 --warmup 3 --no-cpu-baseline --no-extra"
$CMD > gpurun_out/plain.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_C5.csv $CMD > gpurun_out/ncu1.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:mlp_residual_tc -s 4 -c 1 -o gpurun_out/prof_res32 $CMD > gpurun_out/ncu2.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:kl_integrate_tc -s 4 -c 1 -o gpurun_out/prof_int32 $CMD > gpurun_out/ncu3.log 2>&1
ls -la gpurun_out
