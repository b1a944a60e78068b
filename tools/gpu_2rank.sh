#!/bin/bash
# two ranks over NCCL: the default (weak-scaling) bench line and BASELINE configs[3] (strong scaling, sharded pair list)
mkdir -p gpurun_out
T="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
echo "== bench default, 2 ranks"; timeout 900 $T --master-port 29511 bench.py --gpus 2 --steps 10 --warmup 3 2>&1 | tee gpurun_out/bench_2gpu.log | grep '^{' | tail -1 | cut -c1-3000
echo "== bench bates, 2 ranks"; timeout 900 $T --master-port 29512 bench.py --gpus 2 --workload bates --steps 3 --warmup 1 2>&1 | tee gpurun_out/bench_bates_2gpu.log | grep '^{' | tail -1 | cut -c1-2400
echo "== bench bates, 1 rank"; timeout 900 python bench.py --workload bates --steps 3 --warmup 1 2>&1 | tee gpurun_out/bench_bates_1gpu.log | tail -1 | cut -c1-2400
