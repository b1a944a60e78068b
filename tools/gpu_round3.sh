#!/bin/bash
# new 24-warp kernel: debug tile, parity, A/B, full bench with e2e timeline
mkdir -p gpurun_out
echo "== pytest gpu"; timeout 900 python -m pytest tests -m gpu -q -x 2>&1 | tee gpurun_out/pytest_gpu.log | tail -4
echo "== pytest gpu (no cluster)"; IAM_UMMA_NO_CLUSTER=1 timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -q -x 2>&1 | tee gpurun_out/pytest_gpu_nocluster.log | tail -2
B="timeout 600 python bench.py --steps 10 --warmup 3 --no-e2e --no-cpu"
P='import sys,json; d=json.loads(sys.stdin.read()); print(d["value"], d["ms_per_step"], d["roofline"]["kernel_ms_per_launch"], d["roofline"]["frac"], d["clocks"])'
echo "== bench cluster"; $B 2>&1 | tee gpurun_out/bench_new.log | tail -1 | python -c "$P"
echo "== bench no cluster"; IAM_UMMA_NO_CLUSTER=1 $B 2>&1 | tee gpurun_out/bench_nocluster.log | tail -1 | python -c "$P"
echo "== bench no-epilogue"; IAM_UMMA_DEBUG=1 $B 2>&1 | tee gpurun_out/bench_dbg1.log | tail -1 | python -c "$P"
echo "== bench fast-path only"; IAM_UMMA_DEBUG=2 $B 2>&1 | tee gpurun_out/bench_dbg2.log | tail -1 | python -c "$P"
echo "== bench full"; timeout 900 python bench.py 2>&1 | tee gpurun_out/bench_full.log | tail -1 | python -c 'import sys,json; d=json.loads(sys.stdin.read()); print(d["value"], d["roofline"]["frac"], json.dumps(d["e2e"]))'
echo "== ncu full (knn kernel)"
timeout 1200 ncu --set full --clock-control none --import-source on -k regex:knn_umma -s 1 -c 1 -o gpurun_out/knn_umma python bench.py --steps 1 --warmup 1 --frames 60 --no-e2e --no-cpu > gpurun_out/ncu_full.log 2>&1
