#!/bin/bash
# Packed-key epilogue bring-up: parity first, then bench of the default library against the A/B variants in
# imageanalysis_b200/lib/ab_*.so and the profiling ladder (IAM_UMMA_DEBUG) of the default library.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.limit --format=csv > gpurun_out/gpu.txt 2>&1
echo "== pytest gpu parity"; timeout 1200 python -m pytest tests/test_gpu_parity.py -m gpu -q -x 2>&1 | tee gpurun_out/pytest_gpu.log | tail -15
B="timeout 600 python bench.py --steps 10 --warmup 3 --no-e2e --no-cpu"
P='import sys,json; d=json.loads(sys.stdin.read()); r=d["roofline"]; print(d["value"], d["ms_per_step"], r["kernel_ms_per_launch"], r["frac"], r.get("mma_kind"), d["clocks"])'
echo "== bench default"; $B 2>&1 | tee gpurun_out/bench_packed.log | tail -1 | python -c "$P"
for lib in imageanalysis_b200/lib/ab_*.so; do
  [ -f "$lib" ] || continue
  echo "== bench $lib"; IAMATCH_LIB=$PWD/$lib $B 2>&1 | tee gpurun_out/bench_$(basename $lib .so).log | tail -1 | python -c "$P"
done
for m in 1 4 5; do
  echo "== bench default debug mode $m"; IAM_UMMA_DEBUG=$m $B 2>&1 | tee gpurun_out/bench_packed_dbg$m.log | tail -1 | python -c "$P"
done
echo "== bench full (e2e + cpu)"; timeout 900 python bench.py --steps 10 --warmup 3 2>&1 | tee gpurun_out/bench_full.log | tail -1
