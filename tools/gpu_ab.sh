#!/bin/bash
# A/B of kernel builds on one GPU: parity of the default library, then the kNN kernel time of every
# imageanalysis_b200/lib/ab_*.so (built with `make BUILD=../build_x OUT=../lib/ab_x.so EXTRA=-D...`) and of the default.
mkdir -p gpurun_out
B="timeout 600 python bench.py --steps 10 --warmup 3 --no-e2e --no-cpu"
P='import sys,json; d=json.loads(sys.stdin.read()); r=d["roofline"]; print("pairs/s %.0f  ms/step %.3f  knn_ms %.3f  reduce_ms %.3f" % (d["value"], d["ms_per_step"], r["kernel_ms_per_launch"], r["reduce_ms_per_step"]))'
for lib in imageanalysis_b200/lib/libiamatch.so imageanalysis_b200/lib/ab_*.so; do
  [ -f "$lib" ] || continue
  echo "== parity $lib"; IAMATCH_LIB=$PWD/$lib timeout 1200 python -m pytest tests/test_gpu_parity.py -m gpu -q -x 2>&1 | tail -3
  for rep in 1 2; do
    echo "== $lib run $rep"; IAMATCH_LIB=$PWD/$lib $B 2>&1 | tee gpurun_out/ab_$(basename $lib .so)_$rep.log | tail -1 | python -c "$P"
  done
done
${EXTRA_CMD:-true}
