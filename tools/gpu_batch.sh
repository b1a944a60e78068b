bash tools/gpu_ab.sh 2>&1 | grep -v "^\.\.\."
echo "== ncu full ransac"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:ransac_kernel -s 1 -c 1 -f -o gpurun_out/ransac python tools/bench_stages.py --frames 20 --steps 1 --warmup 1 > gpurun_out/ncu_ransac_full.log 2>&1
ls -la gpurun_out/ransac.ncu-rep
