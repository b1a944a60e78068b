echo "== pytest gpu (all)"; timeout 1500 python -m pytest tests -m gpu -q 2>&1 | tail -6
echo "== default bench"; timeout 900 python bench.py 2>&1 | tee gpurun_out/bench_default.log | tail -1 | python -c "
import sys,json; d=json.loads(sys.stdin.read()); e=d['e2e']; r=d['roofline']
print('value %.0f ms/step %.3f e2e %.0f (%.2f ms) knn %.3f ms reduce %.3f ms frac %.3f spot %s orb %s' % (d['value'], d['ms_per_step'], e['value'], e['ms_per_step'], r['kernel_ms_per_launch'], r['reduce_ms_per_step'], r['frac'], d['parity_spot']['status'], d['orb'] and (d['orb']['value'], d['orb']['parity_spot'])))"
for lib in imageanalysis_b200/lib/libiamatch.so imageanalysis_b200/lib/ab_r*.so; do
  echo "== ransac stage with $lib"; IAMATCH_LIB=$PWD/$lib timeout 600 python tools/bench_stages.py --frames 20 --steps 1 --warmup 1 2>&1 | grep -E "ransac_kernel|ORB detect" | python -c "
import sys,json
for l in sys.stdin:
    d=json.loads(l)
    if 'ransac' in d['stage']: print('  ransac ms_per_call %.2f pairs/s %.0f iou_min %.3f' % (d['ms_per_call'], d['pairs_per_s'], d['cpu_baseline']['inlier_iou_vs_gpu_min']))
    else: print('  orb n=%d %.2f ms/frame identical %s cv2 %.1f frames/s' % (d['nfeatures'], d['ms_per_frame'], d['identical_to_cv2'], d['cpu_baseline']['value']))"
done
cp gpurun_out/stages.jsonl gpurun_out/stages_r02.jsonl
