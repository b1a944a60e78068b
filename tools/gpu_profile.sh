#!/bin/bash
# Profiling session for profiles/: full bench line (e2e + cpu), reference arm, ORB, ncu launch list, DRAM traffic of one
# bench-sized launch, one ncu --set full capture of the kNN kernel.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.limit --format=csv > gpurun_out/gpu.txt 2>&1
echo "== pytest gpu (all)"; timeout 1500 python -m pytest tests -m gpu -q 2>&1 | tee gpurun_out/pytest_gpu.log | tail -3
echo "== bench full"; timeout 900 python bench.py --steps 20 --warmup 3 2>&1 | tee gpurun_out/bench_full.log | tail -1 | cut -c1-3000
echo "== bench reference arm"; timeout 600 python bench.py --impl reference --steps 3 --warmup 1 2>&1 | tee gpurun_out/bench_reference.log | tail -1 | cut -c1-800
echo "== bench ORB"; timeout 600 python bench.py --steps 10 --warmup 3 --detector ORB --no-e2e --no-cpu 2>&1 | tee gpurun_out/bench_orb.log | tail -1 | cut -c1-1500
echo "== bench ORB, MMA/TMA pipeline only"; IAM_UMMA_DEBUG=1 timeout 600 python bench.py --steps 10 --warmup 3 --detector ORB --no-e2e --no-cpu 2>&1 | tail -1 | cut -c1-400
for lib in imageanalysis_b200/lib/ab_*.so; do [ -f "$lib" ] && { echo "== bench $lib"; IAMATCH_LIB=$PWD/$lib timeout 600 python bench.py --steps 10 --warmup 3 --no-e2e --no-cpu 2>&1 | tail -1 | cut -c1-400; }; done
echo "== bench fp16 operands"; timeout 600 python bench.py --steps 10 --warmup 3 --engine umma_f16 --no-e2e --no-cpu 2>&1 | tee gpurun_out/bench_f16.log | tail -1 | cut -c1-1500
echo "== stage benchmarks"; timeout 900 python tools/bench_stages.py 2>&1 | tail -2 | cut -c1-1500
echo "== ncu launch list"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"knn_umma_kernel|knn_simt_kernel|metric_reduce_kernel|dedupe_kernel|crosscheck_kernel|finish_dist_kernel|convert_l2_kernel|convert_i8_kernel|convert_hamming_kernel|pack_knn_kernel|gms_kernel|ba_kernel" -c 600 --csv --log-file gpurun_out/launches.csv python bench.py --steps 2 --warmup 1 --frames 40 --no-cpu > gpurun_out/ncu_list.log 2>&1
echo "== ncu dram traffic of one bench-sized launch"
timeout 900 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none -k regex:knn_umma -s 1 -c 1 --csv --log-file gpurun_out/traffic.csv python bench.py --steps 1 --warmup 1 --no-e2e --no-cpu > gpurun_out/ncu_traffic.log 2>&1
tail -3 gpurun_out/traffic.csv
echo "== ncu full (knn kernel)"
timeout 1200 ncu --set full --clock-control none --import-source on -k regex:knn_umma -s 1 -c 1 -f -o gpurun_out/knn_umma python bench.py --steps 1 --warmup 1 --frames 60 --no-e2e --no-cpu > gpurun_out/ncu_full.log 2>&1
echo "== ncu full (knn kernel, ORB)"
timeout 1200 ncu --set full --clock-control none --import-source on -k regex:knn_umma -s 1 -c 1 -f -o gpurun_out/knn_umma_orb python bench.py --steps 1 --warmup 1 --frames 60 --detector ORB --no-e2e --no-cpu > gpurun_out/ncu_full_orb.log 2>&1
echo "== ncu full (ba kernel)"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:ba_kernel -s 2 -c 1 -f -o gpurun_out/ba_kernel python tools/bench_stages.py --frames 20 > gpurun_out/ncu_ba.log 2>&1
ls -la gpurun_out | head -40
