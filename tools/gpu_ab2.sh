#!/bin/bash
# A/B: number of group tests moved to the FMA pipe (build-time switch)
mkdir -p gpurun_out
B="timeout 600 python bench.py --steps 10 --warmup 3 --no-e2e --no-cpu"
P='import sys,json; d=json.loads(sys.stdin.read()); print(d["value"], d["ms_per_step"], d["roofline"]["kernel_ms_per_launch"], d["roofline"]["frac"])'
for G in 0 4 8 6; do
  make -C imageanalysis_b200/csrc clean > /dev/null; make -C imageanalysis_b200/csrc -j16 EXTRA=-DIAM_FMA_GROUPS=$G > gpurun_out/build_g$G.log 2>&1
  echo "== FMA groups $G: parity"; timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "golden or ragged or full_size or extreme" 2>&1 | tail -1
  echo "== FMA groups $G: bench"; $B 2>&1 | tee gpurun_out/bench_g$G.log | tail -1 | python -c "$P"
done
echo "== ncu (6 groups)"; timeout 900 ncu --set full --clock-control none --import-source on -k regex:knn_umma -s 1 -c 1 -o gpurun_out/knn_umma python bench.py --steps 1 --warmup 1 --frames 60 --no-e2e --no-cpu > gpurun_out/ncu_full.log 2>&1
