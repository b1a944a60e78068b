#!/bin/bash
# One GPU session: debug tile -> parity tests -> smoke -> bench.  Every stage has
# its own timeout so a wedged kernel cannot eat the box; logs land in gpurun_out/.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.limit --format=csv > gpurun_out/gpu.txt 2>&1
echo "== debug tile"; timeout 180 python tools/debug_tile.py 2>&1 | tee gpurun_out/debug_tile.log | tail -12
echo "== pytest gpu"; timeout 900 python -m pytest tests -m gpu -q -x 2>&1 | tee gpurun_out/pytest_gpu.log | tail -15
echo "== pytest gpu (all, no -x)"; timeout 900 python -m pytest tests -m gpu -q 2>&1 | tee gpurun_out/pytest_gpu_all.log | tail -25
echo "== smoke"; timeout 300 python __graft_entry__.py --smoke 2>&1 | tee gpurun_out/smoke.log | tail -5
echo "== bench umma"; timeout 900 python bench.py --steps 3 --warmup 3 2>&1 | tee gpurun_out/bench_umma.log | tail -3
echo "== bench simt"; timeout 900 python bench.py --steps 2 --warmup 1 --engine simt --no-e2e --no-cpu 2>&1 | tee gpurun_out/bench_simt.log | tail -3
