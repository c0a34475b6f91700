#!/bin/bash
# 8-GPU session: strong scaling of the default C5 bench (2^24 particles over 8 ranks) and the interacting mean-field system (C4mf, weak)
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29517"
timeout 600 $TR bench.py --gpus 8 --steps 10 --warmup 3 > gpurun_out/bench_C5_8gpu.json 2> gpurun_out/bench_C5_8gpu.err; echo "rc=$?" >> gpurun_out/bench_C5_8gpu.err
timeout 600 $TR bench.py --gpus 8 --steps 10 --warmup 3 --workload C4mf > gpurun_out/bench_C4mf_8gpu.json 2> gpurun_out/bench_C4mf_8gpu.err; echo "rc=$?" >> gpurun_out/bench_C4mf_8gpu.err
tail -n 3 gpurun_out/bench_C5_8gpu.err gpurun_out/bench_C4mf_8gpu.err
