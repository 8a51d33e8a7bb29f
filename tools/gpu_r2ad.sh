#!/bin/bash
# 4-GPU check of the bench (cycle leg over peer memory with ranks that hold no member in the last round; e2e with NUMA binding)
mkdir -p gpurun_out
T=r2ad
timeout 420 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29517 \
  bench.py --gpus 4 --steps 3 --warmup 3 --no-cpu --no-extra --main-deadline 150 > gpurun_out/${T}_bench_n4.json 2> gpurun_out/${T}_bench_n4.err
echo "n4 rc $?"; head -c 300 gpurun_out/${T}_bench_n4.json; echo; tail -3 gpurun_out/${T}_bench_n4.err
