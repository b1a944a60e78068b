"""CPU restatement of cv2.SIFT_create().detectAndCompute(gray, None) -- TEST INFRASTRUCTURE, not product code.

The reference's default detector is SIFT: `detector = cv2.SIFT_create()` (scripts/lib/image.py:236-237) and
`detector.detectAndCompute(scaled, None)` (:324).  The arithmetic lives in OpenCV (features2d/sift.dispatch.cpp,
sift.simd.hpp; pinned `opencv = 4.0.1` in environment.yml, 4.13.0 in this image), which is not vendored under
/root/reference, so this file restates Lowe's published method (IJCV 2004) with OpenCV's defaults (3 layers per
octave, sigma 1.6, contrast threshold 0.04, edge threshold 10, first octave -1) stage by stage.  It was pinned against
live cv2 in the build container (tests/test_sift.py, tests/golden/make_golden_sift.py):

  doubling     cv2.resize(float image, 2x, INTER_LINEAR)                     -- np_up is bit exact
  next octave  cv2.resize(layer 3, 1/2, INTER_NEAREST)                      -- np_half is bit exact
  blur         cv2.GaussianBlur(float image, (0, 0), sigma): separable float filter, ksize = round(8 sigma + 1) | 1,
               reflect-101 borders.  np_blur sums the taps in order; OpenCV's SIMD filter sums them in another order,
               so single pixels differ by <= 1e-4 grey levels (the only stage that is not bit exact)
  extrema, sub-pixel fit, orientation histogram, descriptor: OpenCV's loop order and float expressions, no FMA.
  With cv2's own blur plugged in, 99.7 % of cv2's key points are reproduced to 1e-3 px with descriptors within +-1;
  with np_blur 99.5 %.  The remaining key points sit on a decision boundary (an extremum test, the 0.8 orientation
  peak threshold, a Newton step crossing 0.5) that flips with the last bit of the blurred image; cv2 itself is not
  bit-reproducible across its SIMD dispatch levels there.  The tolerance the tests state follows from this.

Pure-Python loops per key point: use on small images only.
Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg may import this module.
"""
from __future__ import annotations

import math

import numpy as np

NL = 3; SIGMA = 1.6; CONTR = 0.04; EDGE = 10.0; BORDER = 5; MAX_STEPS = 5
ORI_BINS = 36; ORI_SIG = 1.5; ORI_RADIUS = 3 * ORI_SIG; ORI_PEAK = 0.8
DW = 4; DB = 8; D_SCL = 3.0; D_MAG_THR = 0.2; D_INT = 512.0
f32 = np.float32
_P1 = f32(0.9997878412794807 * (180 / math.pi)); _P3 = f32(-0.3258083974640975 * (180 / math.pi))
_P5 = f32(0.1555786518463281 * (180 / math.pi)); _P7 = f32(-0.04432655554792128 * (180 / math.pi))
_EPS = f32(2.220446049250313e-16)


def fast_atan2(y, x):
    y = np.asarray(y, f32); x = np.asarray(x, f32)
    ax, ay = np.abs(x), np.abs(y)
    swap = ax < ay
    num = np.where(swap, ax, ay); den = np.where(swap, ay, ax) + _EPS
    c = (num / den).astype(f32); c2 = c * c
    a = (((_P7 * c2 + _P5) * c2 + _P3) * c2 + _P1) * c
    a = np.where(swap, f32(90.0) - a, a)
    a = np.where(x < 0, f32(180.0) - a, a)
    a = np.where(y < 0, f32(360.0) - a, a)
    return a.astype(f32)

def gauss_pyramid(gray, blur, resize_up, resize_half):
    g = gray.astype(f32)
    sig_diff = math.sqrt(max(SIGMA * SIGMA - 0.5 * 0.5 * 4, 0.01))
    base = blur(resize_up(g), sig_diff)
    n_oct = int(np.rint(math.log(float(min(base.shape))) / math.log(2.0) - 2)) + 1
    sig = [SIGMA]
    k = 2.0 ** (1.0 / NL)
    for i in range(1, NL + 3):
        sp = (k ** (i - 1)) * SIGMA
        st = sp * k
        sig.append(math.sqrt(st * st - sp * sp))
    pyr = []
    for o in range(n_oct):
        for i in range(NL + 3):
            if o == 0 and i == 0:
                pyr.append(base)
            elif i == 0:
                pyr.append(resize_half(pyr[(o - 1) * (NL + 3) + NL]))
            else:
                pyr.append(blur(pyr[-1], sig[i]))
    return pyr, n_oct

def solve3(H, b):
    a = H
    d = a[0][0]*(a[1][1]*a[2][2] - a[2][1]*a[1][2]) - a[0][1]*(a[1][0]*a[2][2] - a[2][0]*a[1][2]) + a[0][2]*(a[1][0]*a[2][1] - a[2][0]*a[1][1])
    d = f32(d)
    if d == 0:
        return [f32(0)] * 3
    d = f32(1) / d
    x0 = d*(b[0]*(a[1][1]*a[2][2] - a[1][2]*a[2][1]) - a[0][1]*(b[1]*a[2][2] - a[1][2]*b[2]) + a[0][2]*(b[1]*a[2][1] - a[1][1]*b[2]))
    x1 = d*(a[0][0]*(b[1]*a[2][2] - a[1][2]*b[2]) - b[0]*(a[1][0]*a[2][2] - a[1][2]*a[2][0]) + a[0][2]*(a[1][0]*b[2] - b[1]*a[2][0]))
    x2 = d*(a[0][0]*(a[1][1]*b[2] - b[1]*a[2][1]) - a[0][1]*(a[1][0]*b[2] - b[1]*a[2][0]) + b[0]*(a[1][0]*a[2][1] - a[1][1]*a[2][0]))
    return [f32(x0), f32(x1), f32(x2)]

def adjust(dog, o, layer, r, c):
    img_scale = f32(1.0 / 255.0); ds = f32(img_scale * f32(0.5)); ss = img_scale; cs = f32(img_scale * f32(0.25))
    xi = xr = xc = f32(0)
    i = 0
    while i < MAX_STEPS:
        idx = o * (NL + 2) + layer
        img, prev, nxt = dog[idx], dog[idx - 1], dog[idx + 1]
        dD = [(img[r, c+1] - img[r, c-1]) * ds, (img[r+1, c] - img[r-1, c]) * ds, (nxt[r, c] - prev[r, c]) * ds]
        v2 = img[r, c] * f32(2)
        dxx = (img[r, c+1] + img[r, c-1] - v2) * ss
        dyy = (img[r+1, c] + img[r-1, c] - v2) * ss
        dss = (nxt[r, c] + prev[r, c] - v2) * ss
        dxy = (img[r+1, c+1] - img[r+1, c-1] - img[r-1, c+1] + img[r-1, c-1]) * cs
        dxs = (nxt[r, c+1] - nxt[r, c-1] - prev[r, c+1] + prev[r, c-1]) * cs
        dys = (nxt[r+1, c] - nxt[r-1, c] - prev[r+1, c] + prev[r-1, c]) * cs
        X = solve3([[dxx, dxy, dxs], [dxy, dyy, dys], [dxs, dys, dss]], dD)
        xi, xr, xc = -X[2], -X[1], -X[0]
        if abs(xi) < 0.5 and abs(xr) < 0.5 and abs(xc) < 0.5:
            break
        if abs(xi) > 2**31 / 3 or abs(xr) > 2**31 / 3 or abs(xc) > 2**31 / 3:
            return None
        c += int(np.rint(xc)); r += int(np.rint(xr)); layer += int(np.rint(xi))
        if layer < 1 or layer > NL or c < BORDER or c >= img.shape[1] - BORDER or r < BORDER or r >= img.shape[0] - BORDER:
            return None
        i += 1
    if i >= MAX_STEPS:
        return None
    idx = o * (NL + 2) + layer
    img, prev, nxt = dog[idx], dog[idx - 1], dog[idx + 1]
    dD = [(img[r, c+1] - img[r, c-1]) * ds, (img[r+1, c] - img[r-1, c]) * ds, (nxt[r, c] - prev[r, c]) * ds]
    t = dD[0] * xc + dD[1] * xr + dD[2] * xi
    contr = img[r, c] * img_scale + t * f32(0.5)
    if abs(contr) * NL < CONTR:
        return None
    v2 = img[r, c] * f32(2)
    dxx = (img[r, c+1] + img[r, c-1] - v2) * ss
    dyy = (img[r+1, c] + img[r-1, c] - v2) * ss
    dxy = (img[r+1, c+1] - img[r+1, c-1] - img[r-1, c+1] + img[r-1, c-1]) * cs
    tr = dxx + dyy; det = dxx * dyy - dxy * dxy
    if det <= 0 or tr * tr * f32(EDGE) >= f32((EDGE + 1) * (EDGE + 1)) * det:
        return None
    kp = dict(x=f32((f32(c) + xc) * f32(1 << o)), y=f32((f32(r) + xr) * f32(1 << o)),
              octave=o + (layer << 8) + (int(np.rint((float(xi) + 0.5) * 255)) << 16),
              size=f32(f32(SIGMA) * f32(math.pow(2.0, float((f32(layer) + xi) / f32(NL)))) * f32(1 << o) * f32(2)),
              response=f32(abs(contr)), r=r, c=c, layer=layer, o=o)
    return kp

def ori_hist(img, px, py, radius, sigma):
    n = ORI_BINS
    expf_scale = f32(-1.0) / (f32(2.0) * f32(sigma) * f32(sigma))
    ys = np.arange(-radius, radius + 1); xs = np.arange(-radius, radius + 1)
    ii, jj = np.meshgrid(ys, xs, indexing='ij')
    y = py + ii; x = px + jj
    ok = (y > 0) & (y < img.shape[0] - 1) & (x > 0) & (x < img.shape[1] - 1)
    y, x, ii, jj = y[ok], x[ok], ii[ok], jj[ok]
    dx = img[y, x + 1] - img[y, x - 1]
    dy = img[y - 1, x] - img[y + 1, x]
    W = np.exp(((ii * ii + jj * jj).astype(f32) * expf_scale).astype(f32)).astype(f32)
    Ori = fast_atan2(dy, dx)
    Mag = np.sqrt(dx * dx + dy * dy).astype(f32)
    bins = np.rint(f32(n / 360.0) * Ori).astype(int)
    bins = np.where(bins >= n, bins - n, bins); bins = np.where(bins < 0, bins + n, bins)
    temp = np.zeros(n, f32)
    wm = (W * Mag).astype(f32)
    for b, v in zip(bins, wm):          # sequential float accumulation, as the reference loop
        temp[b] = temp[b] + v
    t = np.concatenate([temp[-2:], temp, temp[:2]])
    hist = ((t[0:n] + t[4:n+4]) * f32(1/16) + (t[1:n+1] + t[3:n+3]) * f32(4/16) + t[2:n+2] * f32(6/16)).astype(f32)
    return hist, hist.max()

def descriptor(img, ptx, pty, ori, scl):
    d, n = DW, DB
    px, py = int(np.rint(ptx)), int(np.rint(pty))
    cos_t = f32(math.cos(float(f32(ori) * f32(math.pi / 180)))); sin_t = f32(math.sin(float(f32(ori) * f32(math.pi / 180))))
    bins_per_rad = f32(n / 360.0)
    exp_scale = f32(-1.0) / f32(d * d * 0.5)
    hist_width = f32(D_SCL) * f32(scl)
    radius = int(np.rint(hist_width * f32(1.4142135623730951) * f32(d + 1) * f32(0.5)))
    radius = min(radius, int(math.sqrt(float(img.shape[1]) ** 2 + float(img.shape[0]) ** 2)))
    cos_t = cos_t / hist_width; sin_t = sin_t / hist_width
    ii, jj = np.meshgrid(np.arange(-radius, radius + 1), np.arange(-radius, radius + 1), indexing='ij')
    ii = ii.ravel(); jj = jj.ravel()
    c_rot = (jj.astype(f32) * cos_t - ii.astype(f32) * sin_t).astype(f32)
    r_rot = (jj.astype(f32) * sin_t + ii.astype(f32) * cos_t).astype(f32)
    rbin = (r_rot + f32(d // 2) - f32(0.5)).astype(f32); cbin = (c_rot + f32(d // 2) - f32(0.5)).astype(f32)
    r = py + ii; c = px + jj
    ok = (rbin > -1) & (rbin < d) & (cbin > -1) & (cbin < d) & (r > 0) & (r < img.shape[0] - 1) & (c > 0) & (c < img.shape[1] - 1)
    r, c, rbin, cbin, c_rot, r_rot = r[ok], c[ok], rbin[ok], cbin[ok], c_rot[ok], r_rot[ok]
    dx = img[r, c + 1] - img[r, c - 1]; dy = img[r - 1, c] - img[r + 1, c]
    W = np.exp(((c_rot * c_rot + r_rot * r_rot) * exp_scale).astype(f32)).astype(f32)
    Ori = fast_atan2(dy, dx); Mag = np.sqrt(dx * dx + dy * dy).astype(f32)
    hist = np.zeros(((d + 2), (d + 2), (n + 2)), f32).ravel()
    obin = ((Ori - f32(ori)) * bins_per_rad).astype(f32)
    mag = (Mag * W).astype(f32)
    r0 = np.floor(rbin).astype(int); c0 = np.floor(cbin).astype(int); o0 = np.floor(obin).astype(int)
    rb = (rbin - r0).astype(f32); cb = (cbin - c0).astype(f32); ob = (obin - o0).astype(f32)
    o0 = np.where(o0 < 0, o0 + n, o0); o0 = np.where(o0 >= n, o0 - n, o0)
    v_r1 = mag * rb; v_r0 = mag - v_r1
    v11 = v_r1 * cb; v10 = v_r1 - v11; v01 = v_r0 * cb; v00 = v_r0 - v01
    v111 = v11 * ob; v110 = v11 - v111; v101 = v10 * ob; v100 = v10 - v101
    v011 = v01 * ob; v010 = v01 - v011; v001 = v00 * ob; v000 = v00 - v001
    idx = ((r0 + 1) * (d + 2) + c0 + 1) * (n + 2) + o0
    for k in range(len(idx)):
        i0 = idx[k]
        hist[i0] += v000[k]; hist[i0 + 1] += v001[k]; hist[i0 + (n + 2)] += v010[k]; hist[i0 + (n + 3)] += v011[k]
        hist[i0 + (d + 2) * (n + 2)] += v100[k]; hist[i0 + (d + 2) * (n + 2) + 1] += v101[k]
        hist[i0 + (d + 3) * (n + 2)] += v110[k]; hist[i0 + (d + 3) * (n + 2) + 1] += v111[k]
    raw = np.zeros(d * d * n, f32)
    for i in range(d):
        for j in range(d):
            i0 = ((i + 1) * (d + 2) + (j + 1)) * (n + 2)
            hist[i0] += hist[i0 + n]; hist[i0 + 1] += hist[i0 + n + 1]
            raw[(i * d + j) * n:(i * d + j) * n + n] = hist[i0:i0 + n]
    nrm2 = f32(0)
    for v in raw: nrm2 = f32(nrm2 + v * v)
    thr = f32(np.sqrt(nrm2)) * f32(D_MAG_THR)
    raw = np.minimum(raw, thr)
    nrm2 = f32(0)
    for v in raw: nrm2 = f32(nrm2 + v * v)
    s = f32(D_INT) / max(f32(np.sqrt(nrm2)), f32(1.1920929e-07))
    return np.clip(np.rint(raw * s), 0, 255).astype(np.uint8)

def detect_and_compute(gray, blur=None, up=None, half=None):
    blur, up, half = blur or np_blur, up or np_up, half or np_half
    pyr, n_oct = gauss_pyramid(gray, blur, up, half)
    dog = []
    for o in range(n_oct):
        for i in range(NL + 2):
            dog.append(pyr[o * (NL + 3) + i + 1] - pyr[o * (NL + 3) + i])
    thr = int(math.floor(0.5 * CONTR / NL * 255))
    kps = []
    for o in range(n_oct):
        for i in range(1, NL + 1):
            idx = o * (NL + 2) + i
            img, prev, nxt = dog[idx], dog[idx - 1], dog[idx + 1]
            H, W = img.shape
            if H <= 2 * BORDER or W <= 2 * BORDER: continue
            v = img[BORDER:H-BORDER, BORDER:W-BORDER]
            ismax = (np.abs(v) > thr) & (v > 0); ismin = (np.abs(v) > thr) & (v < 0)
            for arr in (prev, img, nxt):
                for dr in (-1, 0, 1):
                    for dc in (-1, 0, 1):
                        nb = arr[BORDER+dr:H-BORDER+dr, BORDER+dc:W-BORDER+dc]
                        ismax &= v >= nb; ismin &= v <= nb
            rs, cs = np.nonzero(ismax | ismin)
            for r, c in zip(rs + BORDER, cs + BORDER):
                kp = adjust(dog, o, i, int(r), int(c))
                if kp is None: continue
                scl_octv = kp['size'] * f32(0.5) / f32(1 << o)
                hist, omax = ori_hist(pyr[o * (NL + 3) + kp['layer']], kp['c'], kp['r'], int(np.rint(f32(ORI_RADIUS) * scl_octv)), f32(ORI_SIG) * scl_octv)
                mag_thr = f32(omax * f32(ORI_PEAK))
                n = ORI_BINS
                for j in range(n):
                    l = j - 1 if j > 0 else n - 1; r2 = j + 1 if j < n - 1 else 0
                    if hist[j] > hist[l] and hist[j] > hist[r2] and hist[j] >= mag_thr:
                        b = f32(j) + f32(0.5) * (hist[l] - hist[r2]) / (hist[l] - f32(2) * hist[j] + hist[r2])
                        b = n + b if b < 0 else (b - n if b >= n else b)
                        ang = f32(360.0) - f32(360.0 / n) * f32(b)
                        if abs(ang - f32(360.0)) < 1.1920929e-07: ang = f32(0)
                        k2 = dict(kp); k2['angle'] = f32(ang)
                        kps.append(k2)
    # removeDuplicatedSorted
    kps.sort(key=lambda k: (k['x'], k['y'], -k['size'], k['angle'], -k['response'], -k['octave']))
    out = []
    for k in kps:
        if out and out[-1]['x'] == k['x'] and out[-1]['y'] == k['y'] and out[-1]['size'] == k['size'] and out[-1]['angle'] == k['angle']:
            continue
        out.append(k)
    des = []
    for k in out:
        o, layer = k['o'], k['layer']
        # first octave -1: published coordinates are halved; descriptors are computed in the octave's own frame
        img = pyr[o * (NL + 3) + layer]
        scale_pub = f32(0.5)
        ptx = f32(k['x'] * scale_pub); pty = f32(k['y'] * scale_pub); size = f32(k['size'] * scale_pub)
        oct_pub = o - 1
        sc = f32(1.0 / (1 << oct_pub)) if oct_pub >= 0 else f32(1 << -oct_pub)
        ang = f32(360.0) - k['angle']
        if abs(ang - f32(360.0)) < 1.1920929e-07: ang = f32(0)
        des.append(descriptor(img, ptx * sc, pty * sc, ang, size * sc * f32(0.5)))
        k['pub'] = (float(ptx), float(pty), float(size), float(k['angle']), float(k['response']), (k['octave'] & ~255) | ((k['octave'] - 1) & 255))
    return out, np.array(des)


# ---- pyramid primitives ----
def gauss_kernel(sigma):
    ks = int(np.rint(sigma * 4 * 2 + 1)) | 1
    x = np.arange(ks, dtype=np.float64) - (ks - 1) / 2
    k = np.exp(-(x * x) / (2.0 * sigma * sigma))
    return (k / k.sum()).astype(f32)

def np_blur(img, sigma):
    k = gauss_kernel(sigma); r = len(k) // 2
    p = np.pad(img, r, mode='reflect')
    H, W = img.shape
    rows = np.zeros((H + 2 * r, W), f32)
    for i in range(len(k)):
        rows += k[i] * p[:, i:i + W]
    out = np.zeros((H, W), f32)
    for i in range(len(k)):
        out += k[i] * rows[i:i + H]
    return out

def np_up(img):
    H, W = img.shape
    def coef(dn, sn):
        fx = (np.arange(dn) + 0.5) * 0.5 - 0.5
        sx = np.floor(fx).astype(int); a = (fx - sx).astype(f32)
        lo = sx < 0; hi = sx >= sn - 1
        sx = np.where(lo, 0, np.where(hi, sn - 1, sx)); a = np.where(lo | hi, f32(0), a)
        return sx, a
    xi, xa = coef(2 * W, W); yi, ya = coef(2 * H, H)
    x1 = np.minimum(xi + 1, W - 1); y1 = np.minimum(yi + 1, H - 1)
    h = img[:, xi] * (f32(1) - xa) + img[:, x1] * xa
    return (h[yi] * (f32(1) - ya)[:, None] + h[y1] * ya[:, None]).astype(f32)

def np_half(img):
    return np.ascontiguousarray(img[0:2 * (img.shape[0] // 2):2, 0:2 * (img.shape[1] // 2):2])


def detect_arrays(gray, **kw):
    """detect_and_compute as arrays: kp [n, 5] f32 (x, y, size, angle, response), octave [n] i32, des [n, 128] u8 -- the
    layout iam_sift_detect returns, in cv2's order."""
    out, des = detect_and_compute(gray, **kw)
    kp = np.array([k['pub'][:5] for k in out], f32).reshape(-1, 5)
    octv = np.array([k['pub'][5] for k in out], np.int32)
    return kp, octv, des.reshape(-1, 128).astype(np.uint8)


def match_keypoints(kp_a, kp_b, tol_px=0.02, tol_size=0.02, tol_angle=0.5):
    """For every key point of a: index of the key point of b at the same position / size / angle, or -1."""
    idx = np.full(len(kp_a), -1, np.int64)
    if len(kp_b) == 0:
        return idx
    order = np.argsort(kp_b[:, 0], kind="stable")
    xs = kp_b[order, 0]
    for i, k in enumerate(kp_a):
        lo, hi = np.searchsorted(xs, k[0] - tol_px), np.searchsorted(xs, k[0] + tol_px, side="right")
        for j in order[lo:hi]:
            q = kp_b[j]
            da = abs(float(k[3]) - float(q[3])); da = min(da, 360.0 - da)
            if abs(k[1] - q[1]) <= tol_px and abs(k[2] - q[2]) <= tol_size and da <= tol_angle:
                idx[i] = j
                break
    return idx
