#!/bin/bash
# one ncu --set full capture of knn_umma (the default library, or IAMATCH_LIB) -> gpurun_out/knn_umma.ncu-rep
mkdir -p gpurun_out
timeout 900 ncu --set full --clock-control none --import-source on -k regex:knn_umma -s 1 -c 1 -f -o gpurun_out/knn_umma python bench.py --steps 1 --warmup 1 --frames 60 --no-e2e --no-cpu > gpurun_out/ncu_full.log 2>&1
tail -2 gpurun_out/ncu_full.log
