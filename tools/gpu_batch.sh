echo "== pytest gpu (orb)"; timeout 900 python -m pytest tests/test_orb.py -m gpu -q -x 2>&1 | tail -15
echo "== pytest gpu (all)"; timeout 1500 python -m pytest tests -m gpu -q 2>&1 | tail -6
echo "== orb timing"; timeout 300 python - <<'PY'
import time, numpy as np, sys
sys.path.insert(0, '.')
from imageanalysis_b200 import detector
rng = np.random.default_rng(0)
img = rng.integers(0, 256, (1459, 2189)).astype(np.uint8)
import cv2
img = cv2.GaussianBlur(img, (0, 0), 1.5)
for n in (5000, 20000):
    detector.orb_detect_and_compute(img, n)
    t0 = time.perf_counter()
    for _ in range(5): r = detector.orb_detect_and_compute(img, n)
    t = (time.perf_counter() - t0) / 5
    orb = cv2.ORB_create(n); orb.detectAndCompute(img, None)
    t0 = time.perf_counter()
    for _ in range(3): k, d = orb.detectAndCompute(img, None)
    tc = (time.perf_counter() - t0) / 3
    print("ORB %d features on 2189x1459: GPU %.1f ms (%d kps), cv2 %.1f ms (%d kps, %d threads)" % (n, t * 1e3, len(r["pt"]), tc * 1e3, len(k), cv2.getNumThreads()))
PY
