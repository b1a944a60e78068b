#!/bin/bash
# One-GPU check: GPU test suite, smoke, the default bench line (configs[1] + ORB leg + parity spot), the reference arm.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.limit --format=csv > gpurun_out/gpu.txt 2>&1
nproc >> gpurun_out/gpu.txt; free -g | head -2 >> gpurun_out/gpu.txt
if [ -z "$SKIP_TESTS" ]; then echo "== pytest gpu"; timeout 1500 python -m pytest tests -m gpu -q -x 2>&1 | tee gpurun_out/pytest_gpu.log | tail -5; fi
echo "== smoke"; timeout 300 python __graft_entry__.py --smoke 2>&1 | tail -2
echo "== bench default"; timeout 900 python bench.py 2>&1 | tee gpurun_out/bench_default.log | tail -1 | cut -c1-6000
echo "== bench reference arm"; timeout 600 python bench.py --impl reference --steps 3 --warmup 1 2>&1 | tee gpurun_out/bench_reference.log | tail -1 | cut -c1-1500
${EXTRA_CMD:-true}
