#!/bin/bash
# Profiling session for profiles/ (round 2): ncu launch list of the default bench command, DRAM traffic of one
# bench-sized kNN launch, and one ncu --set full capture each of the kNN kernel, the metric reduction, the RANSAC
# kernel, the GMS kernel and the ORB kernels.
mkdir -p gpurun_out
K="knn_umma_kernel|knn_simt_kernel|metric_reduce_kernel|dedupe_kernel|crosscheck_kernel|finish_dist_kernel|convert_l2_kernel|convert_i8_kernel|convert_hamming_kernel|pack_knn_kernel|gms_kernel|ba_kernel|ransac_kernel|scan_counts_kernel|pack_tables_kernel"
echo "== ncu launch list (default bench command, short)"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"$K" -c 800 --csv --log-file gpurun_out/launches.csv python bench.py --steps 2 --warmup 1 --frames 40 --no-cpu --no-orb --no-spot > gpurun_out/ncu_list.log 2>&1
echo "== ncu dram traffic of one bench-sized launch"
timeout 900 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none -k regex:knn_umma -s 1 -c 1 --csv --log-file gpurun_out/traffic.csv python bench.py --steps 1 --warmup 1 --no-e2e --no-cpu --no-orb --no-spot > gpurun_out/ncu_traffic.log 2>&1
tail -3 gpurun_out/traffic.csv | cut -c1-300
echo "== ncu full (knn kernel)"
timeout 1200 ncu --set full --clock-control none --import-source on -k regex:knn_umma -s 1 -c 1 -f -o gpurun_out/knn_umma python bench.py --steps 1 --warmup 1 --frames 60 --no-e2e --no-cpu --no-orb --no-spot > gpurun_out/ncu_full.log 2>&1
echo "== ncu full (metric reduction)"
timeout 1200 ncu --set full --clock-control none -k regex:metric_reduce -s 1 -c 1 -f -o gpurun_out/metric_reduce python bench.py --steps 1 --warmup 1 --frames 60 --no-e2e --no-cpu --no-orb --no-spot > gpurun_out/ncu_reduce.log 2>&1
echo "== ncu full (ransac, gms) on the stage benchmarks"
timeout 1200 ncu --set full --clock-control none -k regex:"ransac_kernel|gms_kernel" -s 2 -c 2 -f -o gpurun_out/stages python tools/bench_stages.py --frames 40 --steps 1 --warmup 1 > gpurun_out/ncu_stages_full.log 2>&1
echo "== ncu launch list (orb)"
timeout 600 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -k regex:"resize_kernel|border_kernel|fast_score_kernel|nms_collect_kernel|harris|describe_kernel|blur_" -c 200 --csv --log-file gpurun_out/launches_orb.csv python -c "
import numpy as np, cv2
from imageanalysis_b200 import detector
img = cv2.GaussianBlur(np.random.default_rng(0).integers(0, 256, (1459, 2189)).astype(np.uint8), (0, 0), 1.5)
detector.orb_detect_and_compute(img, 20000)
" > gpurun_out/ncu_orb.log 2>&1
ls -la gpurun_out | head -40
echo "== ncu launch list + full capture (sift): first octave's widest blur, the extrema scan of octave 0, the descriptor kernel"
SK="up2_kernel|blur_rows|blur_cols|half_kernel|extrema_kernel|orientation_kernel|descriptor_kernel|gather_rows_kernel"
timeout 600 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -k regex:"$SK" -c 400 --csv --log-file gpurun_out/launches_sift.csv python tools/sift_profile.py 2189 1459 2 > gpurun_out/ncu_sift.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"descriptor_kernel|orientation_kernel" -c 2 -f -o gpurun_out/sift_keypoint python tools/sift_profile.py 2189 1459 1 > gpurun_out/ncu_sift_full.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"blur_cols_t|blur_rows_t|extrema_kernel" -s 8 -c 3 -f -o gpurun_out/sift_pyramid python tools/sift_profile.py 2189 1459 1 >> gpurun_out/ncu_sift_full.log 2>&1
