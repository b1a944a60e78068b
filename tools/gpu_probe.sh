#!/bin/bash
# host narrowing throughput on the GPU box + two-CTA shape with the packed epilogue
mkdir -p gpurun_out
nproc; lscpu | grep -E "Model name|Socket|Thread|Core|MHz|L3" 
g++ -O3 -o /tmp/probe tools/host_narrow_probe.cpp -lpthread && for t in 4 8 16 32; do /tmp/probe $t | tail -4 | head -1; done
B="timeout 600 python bench.py --steps 10 --warmup 3 --no-e2e --no-cpu"
P='import sys,json; d=json.loads(sys.stdin.read()); r=d["roofline"]; print(d["value"], d["ms_per_step"], r["kernel_ms_per_launch"], r["frac"], r.get("mma_kind"), d["clocks"])'
echo "== tiles/CTA 1 parity"; IAM_UMMA_CTA_TILES=1 timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -q -x 2>&1 | tail -2
echo "== bench tiles/CTA 1"; IAM_UMMA_CTA_TILES=1 $B 2>&1 | tail -1 | python -c "$P"
echo "== bench ORB"; $B --detector ORB 2>&1 | tail -1 | python -c "$P"
