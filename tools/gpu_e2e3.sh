#!/bin/bash
# e2e with / without the short first waves, match_images parity, stage benchmarks (GMS, BA, RANSAC)
mkdir -p gpurun_out
echo "== pytest match_images"; timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q -k "match_images or chunk" 2>&1 | tail -2
P='import sys,json; d=json.loads(sys.stdin.read()); e=d["e2e"]; print(d["value"], e["value"], e["ms_per_step"], e["h2d_bytes_per_step"], e["frames_narrowed_on_host"], e["timeline_ms"])'
B="timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu"
echo "== ramp (default)"; $B 2>&1 | tee gpurun_out/bench_ramp.log | tail -1 | python -c "$P"
echo "== no ramp"; IAM_WAVE_RAMP=0 $B 2>&1 | tail -1 | python -c "$P"
echo "== ramp again"; $B 2>&1 | tail -1 | python -c "$P"
echo "== stage benchmarks"; timeout 900 python tools/bench_stages.py 2>&1 | tail -3 | cut -c1-1200
