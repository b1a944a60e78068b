#!/bin/bash
# end-to-end number (host buffers in, host tables out) with different wave counts
P='import sys,json; d=json.loads(sys.stdin.read()); e=d["e2e"]; print(round(e["value"]), round(e["ms_per_step"],2), e["timeline_ms"])'
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "images or find_matches" 2>&1 | tail -1
for w in ${WAVES:-32 16 48}; do
  echo -n "IAM_MAX_WAVES=$w: "
  IAM_MAX_WAVES=$w timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu 2>&1 | tail -1 | python -c "$P"
done
