#!/bin/bash
# One GPU session: debug tile -> parity tests -> smoke -> bench A/B -> ncu.  Every stage has its own
# timeout so a wedged kernel cannot eat the box; logs land in gpurun_out/.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.limit --format=csv > gpurun_out/gpu.txt 2>&1
echo "== debug tile"; timeout 180 python tools/debug_tile.py 2>&1 | tee gpurun_out/debug_tile.log | tail -12
echo "== pytest gpu"; timeout 900 python -m pytest tests -m gpu -q -x 2>&1 | tee gpurun_out/pytest_gpu.log | tail -25
echo "== pytest gpu (no cluster)"; IAM_UMMA_NO_CLUSTER=1 timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -q -x 2>&1 | tee gpurun_out/pytest_gpu_nocluster.log | tail -5
echo "== smoke"; timeout 300 python __graft_entry__.py --smoke 2>&1 | tee gpurun_out/smoke.log | tail -3
B="timeout 600 python bench.py --steps 10 --warmup 3 --no-e2e --no-cpu"
P='import sys,json; d=json.loads(sys.stdin.read()); print(d["value"], d["ms_per_step"], d["roofline"]["kernel_ms_per_launch"], d["roofline"]["frac"], d["clocks"])'
echo "== bench cluster"; $B 2>&1 | tee gpurun_out/bench_new.log | tail -1 | python -c "$P"
echo "== bench no cluster"; IAM_UMMA_NO_CLUSTER=1 $B 2>&1 | tee gpurun_out/bench_nocluster.log | tail -1 | python -c "$P"
echo "== bench all-pairs (124750 pairs/step)"; timeout 900 python bench.py --steps 2 --warmup 1 --pairs all --no-e2e --no-cpu 2>&1 | tee gpurun_out/bench_allpairs.log | tail -1 | python -c "$P"
echo "== bench reference arm"; timeout 600 python bench.py --impl reference --steps 3 --warmup 1 2>&1 | tee gpurun_out/bench_reference.log | tail -1 | cut -c1-600
echo "== bench no-epilogue (MMA+TMA only), cluster"; IAM_UMMA_DEBUG=1 $B 2>&1 | tee gpurun_out/bench_dbg1.log | tail -1 | python -c "$P"
echo "== bench fast-path only"; IAM_UMMA_DEBUG=2 $B 2>&1 | tee gpurun_out/bench_dbg2.log | tail -1 | python -c "$P"
echo "== bench read-out only"; IAM_UMMA_DEBUG=3 $B 2>&1 | tee gpurun_out/bench_dbg3.log | tail -1 | python -c "$P"
echo "== bench full"; timeout 900 python bench.py --steps 20 --warmup 3 2>&1 | tee gpurun_out/bench_full.log | tail -1
echo "== bench ORB"; timeout 600 python bench.py --steps 10 --warmup 3 --detector ORB --no-e2e 2>&1 | tee gpurun_out/bench_orb.log | tail -1
echo "== ncu launch list"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"knn_umma_kernel|knn_simt_kernel|metric_reduce_kernel|dedupe_kernel|crosscheck_kernel|finish_dist_kernel|convert_l2_kernel|convert_hamming_kernel|pack_knn_kernel" -c 400 --csv --log-file gpurun_out/launches.csv python bench.py --steps 2 --warmup 1 --frames 40 --no-cpu > gpurun_out/ncu_list.log 2>&1
echo "== ncu dram traffic of one bench-sized launch"
timeout 900 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none -k regex:knn_umma -s 1 -c 1 --csv --log-file gpurun_out/traffic.csv python bench.py --steps 1 --warmup 1 --no-e2e --no-cpu > gpurun_out/ncu_traffic.log 2>&1
echo "== ncu full (knn kernel)"
timeout 1200 ncu --set full --clock-control none --import-source on -k regex:knn_umma -s 1 -c 1 -f -o gpurun_out/knn_umma python bench.py --steps 1 --warmup 1 --frames 60 --no-e2e --no-cpu > gpurun_out/ncu_full.log 2>&1
ls -la gpurun_out
