#!/bin/bash
# Round-end check on one GPU: the whole GPU suite, smoke, the default bench line, BASELINE configs[3] (2812 frames), ORB
mkdir -p gpurun_out
echo "== pytest gpu (all)"; timeout 1500 python -m pytest tests -m gpu -q 2>&1 | tee gpurun_out/pytest_gpu.log | tail -6
echo "== smoke"; timeout 300 python __graft_entry__.py --smoke 2>&1 | tail -2
echo "== bench default"; timeout 900 python bench.py 2>&1 | tee gpurun_out/bench_full.log | tail -1 | cut -c1-3200
echo "== bench bates (2812 frames, geotag pairs)"; timeout 900 python bench.py --workload bates --steps 3 --warmup 1 2>&1 | tee gpurun_out/bench_bates_1gpu.log | tail -1 | cut -c1-2400
echo "== bench ORB"; timeout 600 python bench.py --steps 10 --warmup 3 --detector ORB --no-e2e --no-cpu 2>&1 | tee gpurun_out/bench_orb.log | tail -1 | cut -c1-1600
