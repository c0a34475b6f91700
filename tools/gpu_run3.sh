#!/bin/bash
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
python -m pytest tests -m gpu -q -p no:cacheprovider -s > gpurun_out/pytest4_full.log 2>&1
echo "pytest rc=$?" >> gpurun_out/pytest4_full.log
grep -E "vs oracle|FP tensor|boundary sets|KMV |fit vs|passed|failed|FAILED|rc=" gpurun_out/pytest4_full.log > gpurun_out/pytest4.log
timeout 900 python bench.py --steps 5 --warmup 3 > gpurun_out/bench_C5_d.json 2> gpurun_out/bench_C5_d.err
echo "bench rc=$?" >> gpurun_out/bench_C5_d.err
