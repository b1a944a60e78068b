echo "== pytest gpu"; timeout 1500 python -m pytest tests -m gpu -q -x 2>&1 | tail -6
echo "== stage benchmarks (ransac)"; timeout 900 python tools/bench_stages.py --frames 60 --steps 3 --warmup 1 2>&1 | grep ransac | cut -c1-700
echo "== pipeline (configs[4]) 1 GPU"; timeout 1200 python bench.py --workload pipeline --steps 3 --warmup 1 2>&1 | tee gpurun_out/bench_pipeline_n1.log | tail -1 | cut -c1-5000
echo "== default bench"; timeout 900 python bench.py 2>&1 | tee gpurun_out/bench_default.log | tail -1 | cut -c1-1500
