mkdir -p gpurun_out
SK="up2_kernel|blur_rows|blur_cols|half_kernel|extrema_kernel|orientation_kernel|descriptor_kernel|gather_rows_kernel"
timeout 600 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -k regex:"$SK" -c 400 --csv --log-file gpurun_out/launches_sift.csv python tools/sift_profile.py 2189 1459 2 > gpurun_out/ncu_sift.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"descriptor_kernel|orientation_kernel" -c 2 -f -o gpurun_out/sift_keypoint python tools/sift_profile.py 2189 1459 1 > gpurun_out/ncu_sift_full.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"blur_cols_t|blur_rows_t|extrema_kernel" -s 8 -c 3 -f -o gpurun_out/sift_pyramid python tools/sift_profile.py 2189 1459 1 >> gpurun_out/ncu_sift_full.log 2>&1
ls -la gpurun_out | tail -8
