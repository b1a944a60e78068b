#!/bin/bash
# quick check of a kernel change: parity subset + bench
mkdir -p gpurun_out
P='import sys,json; d=json.loads(sys.stdin.read()); print(d["value"], d["ms_per_step"], d["roofline"]["kernel_ms_per_launch"], d["roofline"]["frac"])'
echo "== parity"; timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -q -x 2>&1 | tail -2
echo "== bench"; timeout 600 python bench.py --steps 10 --warmup 3 --no-e2e --no-cpu 2>&1 | tee gpurun_out/bench_quick.log | tail -1 | python -c "$P"
echo "== bench ORB"; timeout 600 python bench.py --steps 10 --warmup 3 --no-e2e --no-cpu --detector ORB 2>&1 | tee gpurun_out/bench_quick_orb.log | tail -1 | python -c "$P"
echo "== ncu"; timeout 900 ncu --set full --clock-control none --import-source on -k regex:knn_umma -s 1 -c 1 -o gpurun_out/knn_umma python bench.py --steps 1 --warmup 1 --frames 60 --no-e2e --no-cpu > gpurun_out/ncu_full.log 2>&1
