#!/bin/bash
# ncu --set full capture (with source) of one knn_umma launch of a 60-frame strip, for the library in $IAMATCH_LIB
# (default: the in-tree one).  Output: gpurun_out/knn_${TAG}.ncu-rep
TAG=${TAG:-cur}
mkdir -p gpurun_out
timeout 1200 ncu --set full --clock-control none --import-source on -k regex:knn_umma -s 1 -c 1 -f -o gpurun_out/knn_$TAG \
  python bench.py --steps 1 --warmup 1 --frames 60 --no-e2e --no-cpu --no-orb --no-spot > gpurun_out/ncu_$TAG.log 2>&1
tail -2 gpurun_out/ncu_$TAG.log | cut -c1-300
ls -la gpurun_out/knn_$TAG.ncu-rep
