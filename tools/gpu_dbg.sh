#!/bin/bash
# kernel time with parts of the epilogue removed (results invalid): where the time of knn_umma goes
P='import sys,json; d=json.loads(sys.stdin.read()); print(d["roofline"]["kernel_ms_per_launch"])'
for m in ${MODES:-0 1 3 2 4}; do
  echo -n "IAM_UMMA_DEBUG=$m  kernel ms: "
  IAM_UMMA_DEBUG=$m timeout 600 python bench.py --steps 10 --warmup 3 --no-e2e --no-cpu 2>&1 | tail -1 | python -c "$P"
done
