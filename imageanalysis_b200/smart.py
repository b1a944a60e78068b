"""Drop-in for the reference's pair-wise side estimators, scripts/lib/smart.py: the surface elevation under an image
pair from the triangulation of its matches, and the yaw error of an image's pose from the similarity transform between
the matched key points.  matcher.find_matches feeds both back into the pose prediction of the 'smart' strategy
(matcher.py:376, :385-386, :987-1005).

The two numeric kernels run on the GPU through the C-ABI:
  triangulate_features      (smart.py:26-63)   cv2.triangulatePoints          -> iam_triangulate_pairs
  find_affine               (smart.py:66-90)   cv2.estimateAffinePartial2D    -> iam_ransac_pairs(IAM_MODEL_AFFINE_PARTIAL)
`surface_estimates(pairs)` does the triangulation statistics of MANY pairs in one launch.  The running estimates live in
the same property-tree records (`/smart/<image>/tri_surface_pairs/...`, `yaw_pairs/...`) and are averaged with the
reference's weights and cut-offs, so a smart.json written by either side loads in the other.  No CPU fallback.
"""
from __future__ import annotations

import json
import os
from math import atan2, pi, sqrt

import numpy as np

try:
    from props import getNode  # type: ignore
except ImportError:  # pragma: no cover
    from .propshim import getNode

from . import _capi

r2d = 180 / pi
d2r = pi / 180
smart_node = getNode("/smart", True)
device = 0
# cv2.estimateAffinePartial2D's defaults (the reference passes none): RANSAC, 3 px, 2000 iterations, confidence 0.99
AFFINE_THRESHOLD, AFFINE_MAX_ITERS, AFFINE_CONFIDENCE = 3.0, 2000, 0.99
_engine = None


def _eng():
    global _engine
    if _engine is None:
        _engine = _capi.Engine(_capi.NORM_L2, 128, device)
    return _engine


def _camera_K():
    cam = getNode('/config/camera', True)
    return np.array([cam.getFloatEnum('K', i) for i in range(9)], np.float64).reshape(3, 3)      # camera.get_K()


def _image_params():
    cam = getNode('/config/camera', True)
    return cam.getInt('width_px'), cam.getInt('height_px')                                       # camera.get_image_params()


def _rodrigues_matrix(rvec):
    r = np.asarray(rvec, np.float64).ravel()
    th = float(np.linalg.norm(r))
    if th < 1e-300:
        return np.eye(3)
    k = r / th
    Kx = np.array([[0, -k[2], k[1]], [k[2], 0, -k[0]], [-k[1], k[0], 0]])
    return np.cos(th) * np.eye(3) + (1 - np.cos(th)) * np.outer(k, k) + np.sin(th) * Kx


def _proj(img):
    """[R | t] of Image.get_proj (image.py:542-553); smart.py:44-50 rebuilds R from the Rodrigues vector."""
    rvec, tvec = img.get_proj()
    return np.concatenate([_rodrigues_matrix(rvec), np.asarray(tvec, np.float64).reshape(3, 1)], axis=1)


def _usable(i1, i2):
    if i1 == i2 or i2.name not in i1.match_list or len(i1.match_list[i2.name]) == 0:
        return False
    for im in (i1, i2):
        if (not im.kp_list or not len(im.kp_list)) and hasattr(im, "load_features"):
            im.load_features()
    return True


def _matched_uv(i1, i2):
    m = np.asarray(i1.match_list[i2.name], np.int64).reshape(-1, 2)
    uv1 = np.array([i1.kp_list[a].pt for a in m[:, 0]], np.float64).reshape(-1, 2)
    uv2 = np.array([i2.kp_list[b].pt for b in m[:, 1]], np.float64).reshape(-1, 2)
    return uv1, uv2


def _normalised(IK, uv):
    return (IK @ np.concatenate([uv, np.ones((len(uv), 1))], axis=1).T)[:2].T


def surface_estimates(pairs):
    """estimate_surface_elevation for a list of (i1, i2) in ONE launch: [(surface_m, std, dist_m) or (None, None, dist_m)]."""
    IK = np.linalg.inv(_camera_K())
    live, x1, x2, p1, p2, off = [], [], [], [], [], [0]
    for n, (i1, i2) in enumerate(pairs):
        if not _usable(i1, i2):
            continue
        uv1, uv2 = _matched_uv(i1, i2)
        live.append(n)
        x1.append(_normalised(IK, uv1))
        x2.append(_normalised(IK, uv2))
        p1.append(_proj(i1).ravel())
        p2.append(_proj(i2).ravel())
        off.append(off[-1] + len(uv1))
    out = []
    for i1, i2 in pairs:
        diff = np.array(i2.get_camera_pose()[0], np.float64) - np.array(i1.get_camera_pose()[0], np.float64)
        out.append([None, None, float(np.linalg.norm(diff))])
    if live:
        _, stats = _eng().triangulate_pairs(np.array(p1), np.array(p2), np.array(off, np.int32), np.concatenate(x1),
                                            np.concatenate(x2), want_points=False)
        for n, st in zip(live, stats):
            out[n][0], out[n][1] = -float(st[0]), float(st[1])     # NED: down is positive, elevation is its negative
    return [tuple(o) for o in out]


def triangulate_features(i1, i2):
    """smart.py:26-63: the matches of the pair triangulated from the two camera poses; [4, N] with a unit last row."""
    if not _usable(i1, i2):
        return None
    IK = np.linalg.inv(_camera_K())
    uv1, uv2 = _matched_uv(i1, i2)
    pts, _ = _eng().triangulate_pairs(_proj(i1).reshape(1, 12), _proj(i2).reshape(1, 12), np.array([0, len(uv1)], np.int32),
                                      _normalised(IK, uv1), _normalised(IK, uv2))
    return np.concatenate([pts.T, np.ones((1, len(pts)))], axis=0)


def find_affine(i1, i2):
    """smart.py:66-90: cv2.estimateAffinePartial2D(uv2, uv1): the similarity that maps image-2 pixels onto image 1."""
    if not _usable(i1, i2):
        return None
    uv1, uv2 = _matched_uv(i1, i2)
    if len(uv1) < 2:
        return None
    _, model, inl = _eng().ransac_pairs(_capi.MODEL_AFFINE_PARTIAL, np.float32(uv2), np.float32(uv1), np.array([0, len(uv1)], np.int32),
                                        None, AFFINE_THRESHOLD, prob=AFFINE_CONFIDENCE, max_iters=AFFINE_MAX_ITERS)
    if inl[0] < 2:
        return None
    return model[0, :2, :].copy()


def decompose_affine(affine):
    """smart.py:94-113: rotation (degrees), translation, signed scales."""
    tx, ty = affine[0][2], affine[1][2]
    a, b, c, d = affine[0][0], affine[0][1], affine[1][0], affine[1][1]
    sx = sqrt(a * a + b * b)
    if a < 0.0:
        sx = -sx
    sy = sqrt(c * c + d * d)
    if d < 0.0:
        sy = -sy
    angle_deg = atan2(-b, a) * 180.0 / pi
    if angle_deg < -180.0:
        angle_deg += 360.0
    if angle_deg > 180.0:
        angle_deg -= 360.0
    return (angle_deg, tx, ty, sx, sy)


def estimate_surface_elevation(i1, i2):
    """smart.py:116-131."""
    return surface_estimates([(i1, i2)])[0]


def estimate_yaw_error(i1, i2):
    """smart.py:139-190: course of image 2's centre seen in image 1 (from the similarity) against the GPS course."""
    affine = find_affine(i1, i2)
    if affine is None:
        return None, None, None, None
    (rot, tx, ty, sx, sy) = decompose_affine(affine)
    weight = abs(ty / tx) if abs(ty) > 0 else abs(tx)
    diff = np.array(i2.get_camera_pose()[0], np.float64) - np.array(i1.get_camera_pose()[0], np.float64)
    dist = np.linalg.norm(diff)
    direction = diff / dist
    crs_gps = 90 - atan2(direction[0], direction[1]) * r2d
    if crs_gps < 0:
        crs_gps += 360
    if crs_gps > 360:
        crs_gps -= 360
    (w, h) = _image_params()
    cx, cy = int(w * 0.5), int(h * 0.5)
    newc = np.asarray(affine, np.float64).dot(np.array([cx, cy, 1.0]))[:2]
    cdiff = [newc[0] - cx, cy - newc[1]]
    crs_aff = 90 - atan2(cdiff[1], cdiff[0]) * r2d
    (_, air_ypr1, _) = i1.get_aircraft_pose()
    crs_fit = air_ypr1[0] + crs_aff
    yaw_error = crs_gps - crs_fit
    if yaw_error < -180:
        yaw_error += 360
    if yaw_error > 180:
        yaw_error -= 360
    return yaw_error, dist, crs_aff, weight


def _weighted(pairs_node, keep):
    total, count = 0, 0
    for child in pairs_node.getChildren():
        node = pairs_node.getChild(child)
        value, weight = keep(node)
        if value is not None:
            total += value * weight
            count += weight
    return total, count


def update_surface_estimate(i1, i2):
    """smart.py:194-245: record the pair's estimate under both images, refresh their weighted averages."""
    avg, std, dist_m = estimate_surface_elevation(i1, i2)
    if avg is None:
        return None, None
    weight = dist_m * dist_m
    for a, b in ((i1, i2), (i2, i1)):
        tri = smart_node.getChild(a.name, True).getChild("tri_surface_pairs", True)
        rec = tri.getChild(b.name, True)
        rec.setFloat("surface_m", float("%.1f" % avg))
        rec.setInt("weight", weight)
        rec.setFloat("stddev", float("%.1f" % std))
        rec.setInt("dist_m", dist_m)
    cutoff_std = 25             # more than this suggests a bad set of matches
    for a in (i1, i2):
        node = smart_node.getChild(a.name, True)
        total, count = _weighted(node.getChild("tri_surface_pairs", True),
                                 lambda r: (r.getFloat("surface_m"), r.getInt("weight")) if r.getFloat("stddev") < cutoff_std else (None, 0))
        if count > 0:
            node.setFloat("tri_surface_m", float("%.1f" % (total / count)))
    return avg, std


def update_yaw_error_estimate(i1, i2):
    """smart.py:249-283."""
    yaw_error, dist, crs_affine, weight = estimate_yaw_error(i1, i2)
    if yaw_error is None:
        return 0
    i1_node = smart_node.getChild(i1.name, True)
    yaw_node = i1_node.getChild("yaw_pairs", True)
    rec = yaw_node.getChild(i2.name, True)
    rec.setFloat("yaw_error", "%.1f" % yaw_error)
    rec.setFloat("dist_m", "%.1f" % dist)
    rec.setFloat("relative_crs", "%.1f" % crs_affine)
    rec.setFloat("weight", "%.1f" % weight)
    total, count = _weighted(yaw_node, lambda r: (r.getFloat("yaw_error"), r.getInt("weight"))
                             if r.getFloat("dist_m") >= 0.5 and abs(r.getFloat("yaw_error")) <= 30 else (None, 0))
    if count > 0:
        i1_node.setFloat("yaw_error", float("%.1f" % (total / count)))
        return total / count
    return 0


def get_yaw_error_estimate(i1):
    node = smart_node.getChild(i1.name, True)
    return node.getFloat("yaw_error") if node.hasChild("yaw_error") else 0.0


def get_surface_estimate(i1, i2):
    """smart.py:294-317: mean of the two images' triangulated surfaces, else of their SRTM values."""
    n1, n2 = smart_node.getChild(i1.name, True), smart_node.getChild(i2.name, True)
    vals = [n.getFloat("tri_surface_m") for n in (n1, n2) if n.hasChild("tri_surface_m")]
    if vals:
        return sum(vals) / len(vals)
    return (n1.getFloat("srtm_surface_m") + n2.getFloat("srtm_surface_m")) * 0.5


def update_srtm_elevations(proj):
    """smart.py:320-325.  The SRTM tiles come from the reference's lib.srtm (network download, out of scope here)."""
    try:
        from . import srtm  # type: ignore
    except ImportError as e:
        raise _capi.IamError("update_srtm_elevations needs the reference's lib.srtm; set /config/matcher/ground_m instead") from e
    for image in proj.image_list:
        ned, ypr, quat = image.get_camera_pose()
        smart_node.getChild(image.name, True).setFloat("srtm_surface_m", float("%.1f" % srtm.ned_interp([ned[0], ned[1]])))


def set_yaw_error_estimates(proj):
    for image in proj.image_list:
        yaw_node = smart_node.getChild(image.name, True).getChild("yaw_pairs", True)
        image.set_aircraft_yaw_error_estimate(yaw_node.getFloat("yaw_error"))


def load(analysis_dir):
    path = os.path.join(analysis_dir, "smart.json")
    try:
        import props_json  # type: ignore
        props_json.load(path, smart_node)
    except ImportError:
        if os.path.exists(path):
            smart_node.from_dict(json.load(open(path)))


def save(analysis_dir):
    path = os.path.join(analysis_dir, "smart.json")
    try:
        import props_json  # type: ignore
        props_json.save(path, smart_node)
    except ImportError:
        with open(path, "w") as f:
            json.dump(smart_node.to_dict(), f, indent=4, sort_keys=True)
