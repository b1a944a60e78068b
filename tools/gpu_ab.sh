#!/bin/bash
# A/B timing of kernel build variants: every imageanalysis_b200/lib/ab_*.so next to the default library.
# Build a variant with:  make -C imageanalysis_b200/csrc BUILD=../build_x OUT=../lib/ab_x.so EXTRA="-D..."
mkdir -p gpurun_out
P='import sys,json; d=json.loads(sys.stdin.read()); print(d["value"], d["ms_per_step"], d["roofline"]["kernel_ms_per_launch"], d["roofline"]["frac"])'
for lib in imageanalysis_b200/lib/libiamatch.so imageanalysis_b200/lib/ab_*.so; do
  [ -f "$lib" ] || continue
  echo "== $lib"
  IAMATCH_LIB=$PWD/$lib timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q -x 2>&1 | tail -1
  for rep in 1 2; do
    IAMATCH_LIB=$PWD/$lib timeout 600 python bench.py --steps 10 --warmup 3 --no-e2e --no-cpu 2>&1 | tail -1 | python -c "$P"
  done
done
