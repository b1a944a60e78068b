#!/bin/bash
# A/B of the B-tile width (build-time switch), then tests + full bench with the e2e timeline
mkdir -p gpurun_out
B="timeout 600 python bench.py --steps 10 --warmup 3 --no-e2e --no-cpu"
P='import sys,json; d=json.loads(sys.stdin.read()); print(d["value"], d["ms_per_step"], d["roofline"]["kernel_ms_per_launch"], d["roofline"]["frac"], d["clocks"])'
for W in 64 96; do
  make -C imageanalysis_b200/csrc clean > /dev/null; make -C imageanalysis_b200/csrc -j16 EXTRA=-DIAM_BROWS=$W > gpurun_out/build_$W.log 2>&1
  echo "== B tile $W columns: parity"; timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q -x 2>&1 | tail -1
  echo "== B tile $W: bench"; $B 2>&1 | tee gpurun_out/bench_w$W.log | tail -1 | python -c "$P"
  echo "== B tile $W: no-epilogue"; IAM_UMMA_DEBUG=1 $B 2>&1 | tail -1 | python -c "$P"
done
echo "== pytest gpu (default build)"; timeout 900 python -m pytest tests -m gpu -q -x 2>&1 | tee gpurun_out/pytest_gpu.log | tail -2
echo "== bench full"; timeout 900 python bench.py 2>&1 | tee gpurun_out/bench_full.log | tail -1 | python -c 'import sys,json; d=json.loads(sys.stdin.read()); print(d["value"], d["roofline"]["frac"], json.dumps(d["e2e"]))'
echo "== ncu full (knn kernel)"
timeout 1200 ncu --set full --clock-control none --import-source on -k regex:knn_umma -s 1 -c 1 -o gpurun_out/knn_umma python bench.py --steps 1 --warmup 1 --frames 60 --no-e2e --no-cpu > gpurun_out/ncu_full.log 2>&1
