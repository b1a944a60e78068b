#!/bin/bash
# host-narrowing A/B: the end-to-end leg by worker order / thread count / SMs reserved for conversion kernels
mkdir -p gpurun_out
echo "== pytest narrowing"; timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "match_images" 2>&1 | tail -2
P='import sys,json; d=json.loads(sys.stdin.read()); e=d["e2e"]; print(d["value"], e["value"], e["ms_per_step"], e["h2d_bytes_per_step"], e["frames_narrowed_on_host"], e["timeline_ms"])'
B="timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu"
echo "== default"; $B 2>&1 | tee gpurun_out/bench_narrow1.log | tail -1 | python -c "$P"
for r in 6 10 12; do
  echo "== reserve $r SMs"; IAM_RESERVE_SMS=$r $B 2>&1 | tail -1 | python -c "$P"
done
echo "== 16 host threads"; IAM_HOST_THREADS=16 $B 2>&1 | tail -1 | python -c "$P"
echo "== 6 host threads (bwd)"; IAM_HOST_THREADS=6 $B 2>&1 | tail -1 | python -c "$P"
