#!/bin/bash
# 2-GPU session: strong scaling of the default C5 bench, the interacting (mean-field) system at 2 ranks
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511"
timeout 600 $TR bench.py --gpus 2 --steps 5 --warmup 3 > gpurun_out/bench_C5_2gpu.json 2> gpurun_out/bench_C5_2gpu.err; echo "rc=$?" >> gpurun_out/bench_C5_2gpu.err
timeout 600 $TR bench.py --gpus 2 --steps 5 --warmup 3 --workload C4mf > gpurun_out/bench_C4mf_2gpu.json 2> gpurun_out/bench_C4mf_2gpu.err; echo "rc=$?" >> gpurun_out/bench_C4mf_2gpu.err
tail -n 3 gpurun_out/bench_C5_2gpu.err gpurun_out/bench_C4mf_2gpu.err
