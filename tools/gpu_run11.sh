#!/bin/bash
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
python -m pytest tests/test_gpu_models.py -m gpu -q -p no:cacheprovider > gpurun_out/pytest11.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest11.log
for wl in C3 C5; do
timeout 600 python bench.py --steps 5 --warmup 3 --workload $wl --model parametric --no-cpu-baseline --no-extra > gpurun_out/bench_${wl}_param.json 2> gpurun_out/bench_${wl}_param.err; echo "rc=$?" >> gpurun_out/bench_${wl}_param.err
done
