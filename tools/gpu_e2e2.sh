#!/bin/bash
# host-narrowing A/B: the end-to-end leg by mode / thread count
mkdir -p gpurun_out
echo "== pytest narrowing"; timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "match_images" 2>&1 | tail -2
P='import sys,json; d=json.loads(sys.stdin.read()); e=d["e2e"]; print(d["value"], e["value"], e["ms_per_step"], e["h2d_bytes_per_step"], e["frames_narrowed_on_host"], e["timeline_ms"])'
B="timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu"
echo "== default"; $B 2>&1 | tee gpurun_out/bench_narrow1.log | tail -1 | python -c "$P"
echo "== always"; IAM_HOST_NARROW=2 $B 2>&1 | tail -1 | python -c "$P"
for t in 8 10; do echo "== $t host threads"; IAM_HOST_THREADS=$t $B 2>&1 | tail -1 | python -c "$P"; done
echo "== 5 host threads forward"; IAM_NARROW_ORDER=fwd IAM_HOST_THREADS=5 $B 2>&1 | tail -1 | python -c "$P"
echo "== 5 host threads backward"; IAM_HOST_THREADS=5 $B 2>&1 | tail -1 | python -c "$P"
