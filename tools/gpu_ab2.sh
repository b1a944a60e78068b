#!/bin/bash
# A/B of kernel build variants: real kernel (mode 0) and MMA/TMA pipeline only (IAM_UMMA_DEBUG=1), parity first.
mkdir -p gpurun_out
B="timeout 600 python bench.py --steps 10 --warmup 3 --no-e2e --no-cpu"
P='import sys,json; d=json.loads(sys.stdin.read()); r=d["roofline"]; print(d["value"], d["ms_per_step"], r["kernel_ms_per_launch"], r["frac"], r.get("mma_kind"))'
for lib in imageanalysis_b200/lib/libiamatch.so imageanalysis_b200/lib/ab_*.so; do
  [ -f "$lib" ] || continue
  n=$(basename $lib .so)
  echo "== $n parity"; IAMATCH_LIB=$PWD/$lib timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -q -x 2>&1 | tail -1
  for m in 0 1 ${EXTRA_MODES}; do
    echo "== $n mode $m"; IAMATCH_LIB=$PWD/$lib IAM_UMMA_DEBUG=$m $B 2>&1 | tee gpurun_out/ab2_${n}_m$m.log | tail -1 | python -c "$P"
  done
done
