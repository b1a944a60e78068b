#!/bin/bash
# A/B of run-time kernel shapes (IAM_UMMA_CTA_TILES) with the default library: parity, then modes 0/1
mkdir -p gpurun_out
B="timeout 600 python bench.py --steps 10 --warmup 3 --no-e2e --no-cpu"
P='import sys,json; d=json.loads(sys.stdin.read()); r=d["roofline"]; print(d["value"], d["ms_per_step"], r["kernel_ms_per_launch"], r["frac"], r.get("mma_kind"))'
for t in 1 2; do
  echo "== tiles/CTA $t parity"; IAM_UMMA_CTA_TILES=$t timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -q -x 2>&1 | tail -3
  for m in 0 1 ${EXTRA_MODES}; do
    echo "== tiles/CTA $t mode $m"; IAM_UMMA_CTA_TILES=$t IAM_UMMA_DEBUG=$m $B 2>&1 | tee gpurun_out/ab3_t${t}_m$m.log | tail -1 | python -c "$P"
  done
done
