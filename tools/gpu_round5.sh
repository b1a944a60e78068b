#!/bin/bash
mkdir -p gpurun_out
B="timeout 600 python bench.py --steps 10 --warmup 3 --no-e2e --no-cpu"
P='import sys,json; d=json.loads(sys.stdin.read()); print(d["value"], d["ms_per_step"], d["roofline"]["kernel_ms_per_launch"], d["roofline"]["frac"])'
for S in 2 1; do
  make -C imageanalysis_b200/csrc clean > /dev/null; make -C imageanalysis_b200/csrc -j16 EXTRA=-DIAM_SHARE_EVERY=$S > gpurun_out/build_s$S.log 2>&1
  echo "== share every $S: bench"; $B 2>&1 | tee gpurun_out/bench_s$S.log | tail -1 | python -c "$P"
done
echo "== pytest gpu (default build)"; timeout 900 python -m pytest tests -m gpu -q -x 2>&1 | tee gpurun_out/pytest_gpu.log | tail -2
echo "== bench full"; timeout 900 python bench.py 2>&1 | tee gpurun_out/bench_full.log | tail -1 | python -c 'import sys,json; d=json.loads(sys.stdin.read()); print(d["value"], d["roofline"]["frac"], json.dumps(d["e2e"]))'
