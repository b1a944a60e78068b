#!/bin/bash
# parity of the default library, then bench of it and of every imageanalysis_b200/lib/ab_*.so variant (twice: noise)
mkdir -p gpurun_out
echo "== parity"; timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -q 2>&1 | tail -3
B="timeout 600 python bench.py --steps 10 --warmup 3 --no-e2e --no-cpu"
P='import sys,json; d=json.loads(sys.stdin.read()); r=d["roofline"]; print(d["value"], d["ms_per_step"], r["kernel_ms_per_launch"], r["frac"])'
for rep in 1; do
for lib in imageanalysis_b200/lib/libiamatch.so imageanalysis_b200/lib/ab_*.so; do
  echo "== $(basename $lib .so) (run $rep)"; IAMATCH_LIB=$PWD/$lib $B 2>&1 | tail -1 | python -c "$P"
done
done
