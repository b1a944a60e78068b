#!/bin/bash
# four ranks over NCCL: the default (weak-scaling) bench line and BASELINE configs[3] (strong scaling)
mkdir -p gpurun_out
T="python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1"
nproc
echo "== bench default, 4 ranks"; timeout 600 $T --master-port 29521 bench.py --gpus 4 --steps 10 --warmup 3 2>&1 | tee gpurun_out/bench_4gpu.log | grep '^{' | tail -1 | cut -c1-3000
echo "== bench bates, 4 ranks"; timeout 600 $T --master-port 29522 bench.py --gpus 4 --workload bates --steps 3 --warmup 1 2>&1 | tee gpurun_out/bench_bates_4gpu.log | grep '^{' | tail -1 | cut -c1-2400
