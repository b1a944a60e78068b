echo "== pytest gpu (all)"; timeout 1500 python -m pytest tests -m gpu -q 2>&1 | tail -6
echo "== default bench"; timeout 900 python bench.py 2>&1 | tee gpurun_out/bench_default.log | tail -1 | python -c "
import sys,json; d=json.loads(sys.stdin.read()); e=d['e2e']; r=d['roofline']
print('value %.0f e2e %.0f (%.2f ms) knn %.3f ms frac %.3f spot %s orb %s' % (d['value'], e['value'], e['ms_per_step'], r['kernel_ms_per_launch'], r['frac'], d['parity_spot']['status'], d['orb'] and (d['orb']['value'], d['orb']['parity_spot'])))"
