#!/bin/bash
# tests + end-to-end bench on 1 GPU, then the 2-rank NCCL path
mkdir -p gpurun_out
echo "== pytest gpu"; timeout 900 python -m pytest tests -m gpu -q -x 2>&1 | tee gpurun_out/pytest_gpu.log | tail -4
echo "== bench full"; timeout 900 python bench.py 2>&1 | tee gpurun_out/bench_full.log | tail -1 | cut -c1-2500
echo "== bench 2 ranks (NCCL)"; timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 10 --warmup 3 2>&1 | tee gpurun_out/bench_2gpu.log | tail -2 | cut -c1-2500
