"""ctypes binding of libiamatch.so (include/iamatch.h).

This is the only place Python touches the native library.  There is no CPU
fallback: if the shared object is missing or no sm_100 device is visible the
constructor raises, and the `matcher` module surfaces that error instead of
quietly calling OpenCV (BASELINE.json north_star: "no CPU fallback").
"""
from __future__ import annotations

import ctypes as C
import os
from typing import Optional, Sequence, Tuple

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
# IAMATCH_LIB selects another build of the same library (A/B timing of kernel variants); default is the in-tree one
LIB_PATH = os.environ.get("IAMATCH_LIB") or os.path.join(_HERE, "lib", "libiamatch.so")

NORM_L2, NORM_HAMMING = 0, 1
DTYPE_U8, DTYPE_F32 = 0, 1
ENGINE_AUTO, ENGINE_UMMA, ENGINE_SIMT, ENGINE_UMMA_F16 = 0, 1, 2, 3
KIND_F16, KIND_F8, KIND_I8 = 0, 1, 2
REDUCE_LOWE, REDUCE_REF_METRIC = 0, 1
MODEL_ESSENTIAL, MODEL_HOMOGRAPHY, MODEL_FUNDAMENTAL, MODEL_AFFINE_PARTIAL = 0, 1, 2, 3

# every symbol include/iamatch.h declares; tests check the library exports them all
EXPORTS = (
    "iam_create", "iam_destroy", "iam_last_error", "iam_abi_version", "iam_set_stream",
    "iam_set_engine", "iam_synchronize", "iam_upload_descriptors",
    "iam_upload_descriptors_device", "iam_upload_keypoint_keys", "iam_upload_keypoints", "iam_upload_keypoints_batch", "iam_gms_filter", "iam_release_descriptors", "iam_num_descriptors",
    "iam_descriptors_exact", "iam_knn_pairs", "iam_match_pairs", "iam_match_pairs_device", "iam_match_images",
    "iam_fetch_tables", "iam_pack_tables_device", "iam_fetch_packed_tables", "iam_ransac_pairs", "iam_ransac_tables", "iam_ba_calib_jacobian", "iam_triangulate_pairs", "iam_orb_detect", "iam_sift_detect", "iam_debug_orb_fast", "iam_set_profiling", "iam_get_timing", "iam_debug_tile",
    "iam_debug_minimal_solver", "iam_ba_setup", "iam_ba_eval", "iam_ba_upload_params", "iam_ba_eval_device",
    "iam_debug_ba_host", "iam_debug_narrow",
)


class MatchParams(C.Structure):
    _fields_ = [
        ("match_ratio", C.c_double),
        ("max_distance", C.c_double),
        ("reduce_mode", C.c_int),
        ("cap", C.c_int),
        ("min_pairs", C.c_int),
        ("cross_check", C.c_int),
        ("dedupe", C.c_int),
        ("gms", C.c_int),
        ("gms_rotation", C.c_int),
        ("gms_scale", C.c_int),
        ("gms_threshold", C.c_double),
        ("width_px", C.c_int),
        ("height_px", C.c_int),
    ]


class Timing(C.Structure):
    _fields_ = [
        ("knn_ms", C.c_float),
        ("reduce_ms", C.c_float),
        ("convert_ms", C.c_float),
        ("knn_launches", C.c_int),
        ("total_launches", C.c_int),
        ("engine_used", C.c_int),
        ("waves", C.c_int),
        ("mma_kind", C.c_int),
        ("host_enqueue_ms", C.c_float),
        ("upload_span_ms", C.c_float),
        ("compute_span_ms", C.c_float),
        ("total_span_ms", C.c_float),
        ("h2d_bytes", C.c_ulonglong),
        ("narrowed_images", C.c_int),
    ]


class IamError(RuntimeError):
    pass


def narrow_host(src):
    """The host-side float32 -> uint8 narrowing of iam_match_images (iam_debug_narrow; needs no GPU).
    Returns (bytes, ok): ok is False when some component is not an integer in 0..255."""
    lib = load_library()
    a = np.ascontiguousarray(src, np.float32)
    out = np.zeros(a.shape, np.uint8)
    bad = lib.iam_debug_narrow(_ptr(a), _ptr(out), a.size)
    if bad < 0:
        raise IamError("iam_debug_narrow failed: %s" % lib.iam_last_error().decode())
    return out, bad == 0


def ba_observation_host(cam7, pt3, uv, K4, dist5, jac: bool = True):
    """The kernel's per-observation function run on the host (iam_debug_ba_host; needs no GPU)."""
    lib = load_library()
    a = [np.ascontiguousarray(v, np.float64) for v in (cam7, pt3, uv, K4, dist5)]
    res = np.zeros(2)
    J = np.zeros((2, 10)) if jac else None
    rc = lib.iam_debug_ba_host(*[_ptr(v) for v in a], _ptr(res), _ptr(J))
    if rc != 0:
        raise IamError("iam_debug_ba_host failed: %s" % lib.iam_last_error().decode())
    return res, J


_lib = None


def load_library(path: Optional[str] = None):
    """dlopen libiamatch.so and declare signatures.  Raises IamError when the
    library has not been built (run `python -m imageanalysis_b200.build`)."""
    global _lib
    if _lib is not None and path is None:
        return _lib
    p = path or LIB_PATH
    if not os.path.exists(p):
        raise IamError(
            f"{p} not found: the CUDA extension is not built. "
            "Run `python -m imageanalysis_b200.build` (needs nvcc); there is no CPU fallback."
        )
    lib = C.CDLL(p)
    vp, ip, i32p, f32p = C.c_void_p, C.c_int, C.POINTER(C.c_int32), C.POINTER(C.c_float)
    lib.iam_create.argtypes = [C.c_int, C.c_int, C.c_int, C.POINTER(vp)]
    lib.iam_destroy.argtypes = [vp]
    lib.iam_last_error.restype = C.c_char_p
    lib.iam_abi_version.restype = C.c_int
    lib.iam_set_stream.argtypes = [vp, vp]
    lib.iam_set_engine.argtypes = [vp, C.c_int]
    lib.iam_synchronize.argtypes = [vp]
    lib.iam_upload_descriptors.argtypes = [vp, C.c_int, vp, C.c_int, C.c_int, C.c_int]
    lib.iam_upload_descriptors_device.argtypes = [vp, C.c_int, vp, C.c_int, C.c_int]
    lib.iam_upload_keypoint_keys.argtypes = [vp, C.c_int, vp, C.c_int]
    lib.iam_upload_keypoints.argtypes = [vp, C.c_int, vp, C.c_int]
    lib.iam_upload_keypoints_batch.argtypes = [vp, C.c_int, vp, vp, vp]
    lib.iam_gms_filter.argtypes = [vp, vp, C.c_int, vp, C.c_int, vp, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_double, vp]
    lib.iam_release_descriptors.argtypes = [vp, C.c_int]
    lib.iam_num_descriptors.argtypes = [vp, C.c_int]
    lib.iam_descriptors_exact.argtypes = [vp, C.c_int]
    lib.iam_knn_pairs.argtypes = [vp, vp, C.c_int, C.c_int, C.c_int, vp, vp, vp, vp]
    lib.iam_match_pairs.argtypes = [vp, vp, C.c_int, C.POINTER(MatchParams), vp, vp, vp, vp]
    lib.iam_match_pairs_device.argtypes = [vp, vp, C.c_int, C.POINTER(MatchParams), C.POINTER(vp), C.POINTER(vp)]
    lib.iam_match_images.argtypes = [vp, C.c_int, vp, vp, vp, C.c_int, vp, vp, C.c_int, C.POINTER(MatchParams), vp, vp]
    lib.iam_fetch_tables.argtypes = [vp, vp, vp]
    lib.iam_pack_tables_device.argtypes = [vp, C.POINTER(vp), C.POINTER(vp), C.POINTER(C.c_longlong)]
    lib.iam_ransac_pairs.argtypes = [vp, C.c_int, vp, vp, vp, C.c_int, vp, C.c_double, C.c_double, C.c_int,
                                     C.c_uint32, vp, vp, vp]
    lib.iam_ransac_tables.argtypes = [vp, C.c_int, vp, C.c_double, C.c_double, C.c_int, C.c_uint32, C.c_int, C.c_int, vp, vp, vp]
    lib.iam_orb_detect.argtypes = [vp, vp, C.c_int, C.c_int, C.c_int, C.c_int, vp, vp, C.POINTER(C.c_int)]
    lib.iam_ba_calib_jacobian.argtypes = [vp, vp, vp, vp]
    lib.iam_fetch_packed_tables.argtypes = [vp, vp, C.c_longlong, vp, C.POINTER(C.c_longlong)]
    lib.iam_triangulate_pairs.argtypes = [vp, C.c_int, vp, vp, vp, vp, vp, vp, vp]
    lib.iam_sift_detect.argtypes = [vp, vp, C.c_int, C.c_int, C.c_int, vp, vp, vp, C.POINTER(C.c_int)]
    lib.iam_debug_orb_fast.argtypes = [vp, vp, C.c_int, C.c_int, vp]
    lib.iam_debug_tile.argtypes = [vp, C.c_int, C.c_int, C.c_int, C.c_int, C.c_uint32, C.c_uint32, C.c_uint32, C.c_int, vp]
    lib.iam_debug_minimal_solver.argtypes = [C.c_int, vp, vp, vp, vp, vp]
    lib.iam_debug_narrow.argtypes = [vp, vp, C.c_size_t]
    lib.iam_ba_setup.argtypes = [vp, C.c_int, C.c_int, C.c_int, vp, vp, vp]
    lib.iam_ba_eval.argtypes = [vp, vp, vp, vp, vp, vp]
    lib.iam_ba_upload_params.argtypes = [vp, vp]
    lib.iam_ba_eval_device.argtypes = [vp, vp, vp, C.c_int, C.POINTER(vp), C.POINTER(vp)]
    lib.iam_debug_ba_host.argtypes = [vp, vp, vp, vp, vp, vp, vp]
    lib.iam_set_profiling.argtypes = [vp, C.c_int]
    lib.iam_get_timing.argtypes = [vp, C.POINTER(Timing)]
    for name in EXPORTS:
        fn = getattr(lib, name)
        if name not in ("iam_last_error",):
            fn.restype = C.c_int
    if path is None:
        _lib = lib
    return lib


def minimal_solver_host(model: int, x1, y1, x2, y2) -> np.ndarray:
    """Host run of the RANSAC minimal solver (no GPU needed); returns [n_models, 3, 3]."""
    lib = load_library()
    arrs = [np.ascontiguousarray(a, np.float32) for a in (x1, y1, x2, y2)]
    out = np.zeros((10, 9), np.float32)
    n = lib.iam_debug_minimal_solver(model, *[a.ctypes.data_as(C.c_void_p) for a in arrs], out.ctypes.data_as(C.c_void_p))
    if n < 0:
        raise IamError("minimal solver failed: %s" % lib.iam_last_error().decode())
    return out[:n].reshape(n, 3, 3)


def _ptr(a: Optional[np.ndarray]):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


class Engine:
    """One matching context on one GPU (wraps iam_ctx)."""

    def __init__(self, norm: int, desc_bytes: int, device: int = 0):
        self._lib = load_library()
        self._h = C.c_void_p()
        self.norm = norm
        self.desc_bytes = desc_bytes
        self.device = device
        rc = self._lib.iam_create(device, norm, desc_bytes, C.byref(self._h))
        if rc != 0:
            self._h = C.c_void_p()
            raise IamError(f"iam_create failed ({rc}): {self._lib.iam_last_error().decode()}")

    # -- plumbing -----------------------------------------------------------
    def _check(self, rc: int, what: str):
        if rc != 0:
            raise IamError(f"{what} failed ({rc}): {self._lib.iam_last_error().decode()}")

    def close(self):
        if getattr(self, "_h", None) and self._h.value:
            self._lib.iam_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def set_stream(self, cuda_stream: Optional[int]):
        self._check(self._lib.iam_set_stream(self._h, C.c_void_p(cuda_stream or 0)), "iam_set_stream")

    def set_engine(self, engine: int):
        self._check(self._lib.iam_set_engine(self._h, engine), "iam_set_engine")

    def synchronize(self):
        self._check(self._lib.iam_synchronize(self._h), "iam_synchronize")

    def set_profiling(self, on: bool):
        self._check(self._lib.iam_set_profiling(self._h, int(on)), "iam_set_profiling")

    def timing(self) -> Timing:
        t = Timing()
        self._check(self._lib.iam_get_timing(self._h, C.byref(t)), "iam_get_timing")
        return t

    # -- descriptors --------------------------------------------------------
    def upload(self, image_id: int, des: np.ndarray, pinned: bool = False):
        """des: [N, D] float32 (SIFT) or uint8 (SIFT-as-u8 / ORB)."""
        des = np.ascontiguousarray(des)
        if des.ndim != 2 or des.shape[1] != self.desc_bytes:
            raise IamError(f"descriptor array must be [N,{self.desc_bytes}], got {des.shape}")
        if des.dtype == np.uint8:
            dt = DTYPE_U8
        elif des.dtype == np.float32:
            dt = DTYPE_F32
        else:
            raise IamError(f"descriptor dtype {des.dtype} unsupported (uint8 / float32)")
        self._check(self._lib.iam_upload_descriptors(self._h, image_id, _ptr(des), des.shape[0], dt, int(pinned)),
                    "iam_upload_descriptors")

    def upload_device(self, image_id: int, dptr: int, n: int, dtype: int):
        self._check(self._lib.iam_upload_descriptors_device(self._h, image_id, C.c_void_p(dptr), n, dtype),
                    "iam_upload_descriptors_device")

    def upload_keypoint_keys(self, image_id: int, keys: np.ndarray):
        keys = np.ascontiguousarray(keys, np.int32)
        self._check(self._lib.iam_upload_keypoint_keys(self._h, image_id, _ptr(keys), keys.shape[0]),
                    "iam_upload_keypoint_keys")

    def upload_keypoints(self, image_id: int, xy: np.ndarray):
        """xy: [N, 2] float32 pixel coordinates (cv2.KeyPoint.pt) for the GMS filter."""
        xy = np.ascontiguousarray(xy, np.float32).reshape(-1, 2)
        self._check(self._lib.iam_upload_keypoints(self._h, image_id, _ptr(xy), xy.shape[0]), "iam_upload_keypoints")

    def upload_keypoints_batch(self, image_ids, xys):
        """upload_keypoints for many images with one synchronisation (iam_upload_keypoints_batch)."""
        ids = np.ascontiguousarray(image_ids, np.int32)
        arrs = [np.ascontiguousarray(a, np.float32).reshape(-1, 2) for a in xys]
        if len(arrs) != len(ids):
            raise IamError("image_ids and xys differ in length")
        n = len(arrs)
        ptrs = (C.c_void_p * max(n, 1))(*[a.ctypes.data for a in arrs])
        counts = np.ascontiguousarray([a.shape[0] for a in arrs], np.int32)
        self._check(self._lib.iam_upload_keypoints_batch(self._h, n, _ptr(ids), ptrs, _ptr(counts)), "iam_upload_keypoints_batch")

    def gms_filter(self, xy1, xy2, matches, size, with_rotation=True, with_scale=False, threshold_factor=5.0,
                   archive_wrap=False):
        """Inlier mask of cv2.xfeatures2d.matchGMS over `matches` ([[queryIdx, trainIdx], ...]); matcher.py:285."""
        xy1 = np.ascontiguousarray(xy1, np.float32).reshape(-1, 2)
        xy2 = np.ascontiguousarray(xy2, np.float32).reshape(-1, 2)
        m = np.ascontiguousarray(matches, np.int32).reshape(-1, 2)
        mask = np.zeros((m.shape[0],), np.uint8)
        self._check(self._lib.iam_gms_filter(self._h, _ptr(xy1), xy1.shape[0], _ptr(xy2), xy2.shape[0], _ptr(m), m.shape[0],
                                             int(size[0]), int(size[1]), int(with_rotation), int(bool(with_scale)) | (2 if archive_wrap else 0),
                                             float(threshold_factor), _ptr(mask)), "iam_gms_filter")
        return mask.astype(bool)

    def release(self, image_id: int):
        self._check(self._lib.iam_release_descriptors(self._h, image_id), "iam_release_descriptors")

    def num_descriptors(self, image_id: int) -> int:
        return self._lib.iam_num_descriptors(self._h, image_id)

    def descriptors_exact(self, image_id: int) -> int:
        return self._lib.iam_descriptors_exact(self._h, image_id)

    # -- kNN ------------------------------------------------------------------
    def knn_pairs(self, pairs: Sequence[Tuple[int, int]], k: int, n_stride: int, reverse: bool = True):
        """Returns (idx_fwd, dist_fwd, idx_rev, dist_rev); arrays are [P, n_stride, k]
        with -1 / NaN in rows beyond an image's descriptor count."""
        pr = np.ascontiguousarray(np.asarray(pairs, dtype=np.int32).reshape(-1, 2))
        P = pr.shape[0]
        idx_f = np.full((P, n_stride, k), -1, np.int32)
        dst_f = np.full((P, n_stride, k), np.nan, np.float32)
        idx_r = np.full((P, n_stride, k), -1, np.int32) if reverse else None
        dst_r = np.full((P, n_stride, k), np.nan, np.float32) if reverse else None
        self._check(self._lib.iam_knn_pairs(self._h, _ptr(pr), P, k, n_stride, _ptr(idx_f), _ptr(dst_f), _ptr(idx_r),
                                            _ptr(dst_r)), "iam_knn_pairs")
        return idx_f, dst_f, idx_r, dst_r

    # -- full match -------------------------------------------------------------
    @staticmethod
    def make_params(match_ratio=0.75, max_distance=270.0, reduce_mode=REDUCE_REF_METRIC, cap=2000, min_pairs=25,
                    cross_check=True, dedupe=False, gms=False, gms_rotation=True, gms_scale=False, gms_threshold=5.0,
                    size=(0, 0)) -> MatchParams:
        p = MatchParams()
        p.match_ratio = float(match_ratio)
        p.max_distance = float(max_distance)
        p.reduce_mode = int(reduce_mode)
        p.cap = int(cap)
        p.min_pairs = int(min_pairs)
        p.cross_check = int(bool(cross_check))
        p.dedupe = int(bool(dedupe))
        p.gms = int(gms) if isinstance(gms, int) and not isinstance(gms, bool) else int(bool(gms))   # 2: archive-Python last-half-cell rule
        p.gms_rotation = int(bool(gms_rotation))
        p.gms_scale = int(bool(gms_scale))
        p.gms_threshold = float(gms_threshold)
        p.width_px, p.height_px = int(size[0]), int(size[1])
        return p

    def match_pairs(self, pairs, params: MatchParams, want_reverse: bool = False):
        pr = np.ascontiguousarray(np.asarray(pairs, dtype=np.int32).reshape(-1, 2))
        P = pr.shape[0]
        table = np.empty((P, params.cap, 2), np.int32)
        count = np.zeros((P,), np.int32)
        rtable = np.empty((P, params.cap, 2), np.int32) if want_reverse else None
        rcount = np.zeros((P,), np.int32) if want_reverse else None
        self._check(self._lib.iam_match_pairs(self._h, _ptr(pr), P, C.byref(params), _ptr(table), _ptr(count),
                                              _ptr(rtable), _ptr(rcount)), "iam_match_pairs")
        if want_reverse:
            return table, count, rtable, rcount
        return table, count

    def match_images(self, image_ids, arrays, pairs, params: MatchParams, keys=None, out=None):
        """Upload + match in one call with PCIe/compute overlap (iam_match_images).
        arrays[i]: [N_i, D] float32 or uint8 descriptors of image image_ids[i] (all the same dtype);
        keys[i]: optional int32 [N_i] keypoint position ids;
        out: optional (table [P, cap, 2] int32, count [P] int32) C-contiguous arrays to receive the result --
        page-locked ones make the wave-by-wave download truly asynchronous."""
        ids = np.ascontiguousarray(image_ids, np.int32)
        arrs = [np.ascontiguousarray(a) for a in arrays]
        if len(arrs) != len(ids):
            raise IamError("image_ids and arrays differ in length")
        dts = {a.dtype for a in arrs}
        if len(dts) > 1 or (arrs and arrs[0].dtype not in (np.uint8, np.float32)):
            raise IamError("descriptor arrays must all be uint8 or all float32")
        dt = DTYPE_U8 if (not arrs or arrs[0].dtype == np.uint8) else DTYPE_F32
        for a in arrs:
            if a.ndim != 2 or a.shape[1] != self.desc_bytes:
                raise IamError(f"descriptor array must be [N,{self.desc_bytes}], got {a.shape}")
        n = len(arrs)
        ptrs = (C.c_void_p * max(n, 1))(*[a.ctypes.data for a in arrs])
        counts = np.ascontiguousarray([a.shape[0] for a in arrs], np.int32)
        kptrs = None
        karrs = None
        if keys is not None:
            karrs = [None if k is None else np.ascontiguousarray(k, np.int32) for k in keys]
            kptrs = (C.c_void_p * max(n, 1))(*[None if k is None else k.ctypes.data for k in karrs])
        pr = np.ascontiguousarray(np.asarray(pairs, dtype=np.int32).reshape(-1, 2))
        P = pr.shape[0]
        if out is not None:
            table, count = out
            if (table.dtype != np.int32 or count.dtype != np.int32 or table.shape != (P, params.cap, 2)
                    or count.shape != (P,) or not table.flags.c_contiguous or not count.flags.c_contiguous):
                raise IamError("out must be (int32 [P, cap, 2], int32 [P]) C-contiguous arrays")
        else:
            table = np.empty((P, params.cap, 2), np.int32)
            count = np.zeros((P,), np.int32)
        self._check(self._lib.iam_match_images(self._h, n, _ptr(ids), ptrs, _ptr(counts), dt, kptrs, _ptr(pr), P,
                                               C.byref(params), _ptr(table), _ptr(count)), "iam_match_images")
        return table, count

    def match_pairs_device(self, pairs: np.ndarray, params: MatchParams) -> Tuple[int, int]:
        """Enqueue only; results stay on the device.  Returns (d_table, d_count) raw pointers."""
        pr = np.ascontiguousarray(np.asarray(pairs, dtype=np.int32).reshape(-1, 2))
        dt, dc = C.c_void_p(), C.c_void_p()
        self._check(self._lib.iam_match_pairs_device(self._h, _ptr(pr), pr.shape[0], C.byref(params), C.byref(dt),
                                                     C.byref(dc)), "iam_match_pairs_device")
        return dt.value or 0, dc.value or 0

    def fetch_tables(self, n_pairs: int, cap: int):
        table = np.empty((n_pairs, cap, 2), np.int32)
        count = np.zeros((n_pairs,), np.int32)
        self._check(self._lib.iam_fetch_tables(self._h, _ptr(table), _ptr(count)), "iam_fetch_tables")
        return table, count

    def fetch_packed_tables(self, n_pairs: int, rows_out: Optional[np.ndarray] = None, offsets_out: Optional[np.ndarray] = None):
        """The last match call's tables in compact (CSR) form on the host (iam_fetch_packed_tables): (rows [total, 2],
        offsets [n_pairs + 1]); pair p owns rows[offsets[p]:offsets[p + 1]].  rows_out / offsets_out: optional
        caller-owned int32 arrays (page-locked ones make the download asynchronous inside the call)."""
        off = offsets_out if offsets_out is not None else np.empty((n_pairs + 1,), np.int32)
        tot = C.c_longlong(0)
        if rows_out is None:
            dr, do, total = self.pack_tables_device()
            rows_out = np.empty((max(total, 1), 2), np.int32)
        if rows_out.dtype != np.int32 or off.dtype != np.int32 or not rows_out.flags.c_contiguous or off.shape[0] < n_pairs + 1:
            raise IamError("fetch_packed_tables wants C-contiguous int32 arrays (rows [cap, 2], offsets [n_pairs + 1])")
        self._check(self._lib.iam_fetch_packed_tables(self._h, _ptr(rows_out), rows_out.shape[0], _ptr(off), C.byref(tot)),
                    "iam_fetch_packed_tables")
        return rows_out[:tot.value], off[:n_pairs + 1]

    def pack_tables_device(self) -> Tuple[int, int, int]:
        """Compact (CSR) form of the last match call's tables, left on the device (iam_pack_tables_device):
        returns (d_rows, d_offsets, total) -- raw pointers to int32 [total, 2] and int32 [n_pairs + 1]."""
        dr, do, tot = C.c_void_p(), C.c_void_p(), C.c_longlong()
        self._check(self._lib.iam_pack_tables_device(self._h, C.byref(dr), C.byref(do), C.byref(tot)),
                    "iam_pack_tables_device")
        return dr.value or 0, do.value or 0, int(tot.value)

    def debug_tile(self, q_id, t_id, q_tile=0, t_tile=0, lbo=128, sbo=2304, kstep_bytes=256, ksteps=9):
        out = np.zeros((128, 128), np.float32)
        self._check(self._lib.iam_debug_tile(self._h, q_id, t_id, q_tile, t_tile, lbo, sbo, kstep_bytes, ksteps,
                                             _ptr(out)), "iam_debug_tile")
        return out

    # -- bundle adjustment ------------------------------------------------------
    def ba_setup(self, n_cam: int, n_pts: int, cam_idx, pt_idx, obs_uv):
        """Problem structure of Optimizer.fun (optimizer.py:396-404): observation i = pixel obs_uv[i] of point
        pt_idx[i] in camera cam_idx[i]."""
        ci = np.ascontiguousarray(cam_idx, np.int32)
        pi = np.ascontiguousarray(pt_idx, np.int32)
        uv = np.ascontiguousarray(obs_uv, np.float64).reshape(-1, 2)
        if not (len(ci) == len(pi) == len(uv)):
            raise IamError("cam_idx, pt_idx and obs_uv differ in length")
        self._check(self._lib.iam_ba_setup(self._h, int(n_cam), int(n_pts), len(ci), _ptr(ci), _ptr(pi), _ptr(uv)),
                    "iam_ba_setup")
        self._ba_shape = (int(n_cam), int(n_pts), len(ci))

    def ba_eval(self, params, K4, dist5, jac: bool = False, out=None):
        """residual [2*n_obs] (and the Jacobian blocks [n_obs, 2, 10]) at `params` (cameras then points).
        out: optional caller-owned (residual, jacobian) float64 arrays, e.g. page-locked ones."""
        n_cam, n_pts, n_obs = self._ba_shape
        p = np.ascontiguousarray(params, np.float64)
        if p.size < n_cam * 7 + n_pts * 3:
            raise IamError("parameter vector shorter than n_cam*7 + n_pts*3")
        k4 = np.ascontiguousarray(K4, np.float64)
        d5 = np.ascontiguousarray(dist5, np.float64)
        if out is not None:
            res, J = out
            if (res.dtype != np.float64 or res.size != 2 * n_obs or not res.flags.c_contiguous or
                    (jac and (J is None or J.dtype != np.float64 or J.size != n_obs * 20 or not J.flags.c_contiguous))):
                raise IamError("out must be (float64 [2 n_obs], float64 [n_obs, 2, 10]) C-contiguous arrays")
            if not jac:
                J = None
        else:
            res = np.empty(2 * n_obs, np.float64)
            J = np.empty((n_obs, 2, 10), np.float64) if jac else None
        self._check(self._lib.iam_ba_eval(self._h, _ptr(p), _ptr(k4), _ptr(d5), _ptr(res), _ptr(J)), "iam_ba_eval")
        return (res, J) if jac else res

    def ba_calib_jacobian(self, K4, dist5):
        """[n_obs, 2, 8] = d residual / d (f, cu, cv, k1, k2, p1, p2, k3) at the parameters last uploaded."""
        n_obs = self._ba_shape[2]
        k4 = np.ascontiguousarray(K4, np.float64)
        d5 = np.ascontiguousarray(dist5, np.float64)
        J = np.empty((n_obs, 2, 8), np.float64)
        self._check(self._lib.iam_ba_calib_jacobian(self._h, _ptr(k4), _ptr(d5), _ptr(J)), "iam_ba_calib_jacobian")
        return J

    def ba_upload_params(self, params):
        p = np.ascontiguousarray(params, np.float64)
        self._check(self._lib.iam_ba_upload_params(self._h, _ptr(p)), "iam_ba_upload_params")

    def ba_eval_device(self, K4, dist5, jac: bool = True):
        k4 = np.ascontiguousarray(K4, np.float64)
        d5 = np.ascontiguousarray(dist5, np.float64)
        dr, dj = C.c_void_p(), C.c_void_p()
        self._check(self._lib.iam_ba_eval_device(self._h, _ptr(k4), _ptr(d5), int(jac), C.byref(dr), C.byref(dj)),
                    "iam_ba_eval_device")
        return dr.value or 0, dj.value or 0

    # -- RANSAC -----------------------------------------------------------------
    def ransac_pairs(self, model: int, pts1: np.ndarray, pts2: np.ndarray, offsets: np.ndarray, K: Optional[np.ndarray],
                     threshold: float, prob: float = 0.999, max_iters: int = 1000, seed: int = 0):
        pts1 = np.ascontiguousarray(pts1, np.float32).reshape(-1, 2)
        pts2 = np.ascontiguousarray(pts2, np.float32).reshape(-1, 2)
        off = np.ascontiguousarray(offsets, np.int32)
        P = off.shape[0] - 1
        Kc = None if K is None else np.ascontiguousarray(K, np.float64).reshape(3, 3)
        mask = np.zeros((pts1.shape[0],), np.uint8)
        model_out = np.zeros((P, 9), np.float64)
        inl = np.zeros((P,), np.int32)
        self._check(self._lib.iam_ransac_pairs(self._h, model, _ptr(pts1), _ptr(pts2), _ptr(off), P, _ptr(Kc),
                                               float(threshold), float(prob), int(max_iters), int(seed), _ptr(mask),
                                               _ptr(model_out), _ptr(inl)), "iam_ransac_pairs")
        return mask, model_out.reshape(P, 3, 3), inl


    def ransac_tables(self, model: int, K: Optional[np.ndarray], threshold: float, n_pairs: int, cap: int, min_pairs: int = 25,
                      compact: bool = True, prob: float = 0.999, max_iters: int = 1000, seed: int = 0, want_mask: bool = False,
                      want_model: bool = True, host_outputs: bool = True):
        """filter_by_transform for every pair of the last match call, on the device tables (iam_ransac_tables).
        Returns (mask [P, cap] or None, models [P, 3, 3] or None, inliers [P])."""
        Kc = None if K is None else np.ascontiguousarray(K, np.float64).reshape(3, 3)
        mask = np.zeros((n_pairs, cap), np.uint8) if want_mask else None
        models = np.zeros((n_pairs, 9), np.float64) if want_model else None
        inl = np.zeros((n_pairs,), np.int32) if host_outputs else None   # all outputs None: the call only enqueues
        self._check(self._lib.iam_ransac_tables(self._h, model, _ptr(Kc), float(threshold), float(prob), int(max_iters),
                                                int(seed), int(min_pairs), int(compact), _ptr(mask), _ptr(models), _ptr(inl)),
                    "iam_ransac_tables")
        return mask, (models.reshape(n_pairs, 3, 3) if want_model else None), inl


    def triangulate_pairs(self, proj1, proj2, offsets, x1, x2, want_points: bool = True):
        """cv2.triangulatePoints for many image pairs in one launch (iam_triangulate_pairs; smart.py:26-63, :116-131).
        proj1/proj2 [P, 12] float64, offsets [P + 1], x1/x2 [total, 2] normalised image coordinates.
        Returns (points [total, 3] float64 or None, stats [P, 2] = mean, std of Z)."""
        p1 = np.ascontiguousarray(proj1, np.float64).reshape(-1, 12)
        p2 = np.ascontiguousarray(proj2, np.float64).reshape(-1, 12)
        off = np.ascontiguousarray(offsets, np.int32)
        a = np.ascontiguousarray(x1, np.float64).reshape(-1, 2)
        b = np.ascontiguousarray(x2, np.float64).reshape(-1, 2)
        P = off.shape[0] - 1
        if p1.shape[0] != P or p2.shape[0] != P or a.shape != b.shape or (P > 0 and a.shape[0] != off[-1]):
            raise IamError("triangulate_pairs: inconsistent shapes")
        pts = np.zeros((a.shape[0], 3), np.float64) if want_points else None
        stats = np.zeros((P, 2), np.float64)
        self._check(self._lib.iam_triangulate_pairs(self._h, P, _ptr(p1), _ptr(p2), _ptr(off), _ptr(a), _ptr(b), _ptr(pts),
                                                    _ptr(stats)), "iam_triangulate_pairs")
        return pts, stats

    def orb_detect(self, gray: np.ndarray, nfeatures: int):
        """cv2.ORB_create(nfeatures).detectAndCompute(gray, None) on the GPU (iam_orb_detect).
        Returns (kp [n, 6] float32 = x, y, size, angle, response, octave; des [n, 32] uint8)."""
        g = np.ascontiguousarray(gray, np.uint8)
        if g.ndim != 2:
            raise IamError("orb_detect wants a 2-D uint8 image")
        cap = int(nfeatures) + 1024
        kp = np.zeros((cap, 6), np.float32)
        des = np.zeros((cap, 32), np.uint8)
        n = C.c_int(0)
        self._check(self._lib.iam_orb_detect(self._h, _ptr(g), g.shape[1], g.shape[0], int(nfeatures), cap, _ptr(kp), _ptr(des),
                                             C.byref(n)), "iam_orb_detect")
        return kp[:n.value].copy(), des[:n.value].copy()

    def sift_detect(self, gray: np.ndarray, max_out: int = 0):
        """cv2.SIFT_create().detectAndCompute(gray, None) on the GPU (iam_sift_detect).
        Returns (kp [n, 5] float32 = x, y, size, angle, response; octave [n] int32; des [n, 128] uint8)."""
        g = np.ascontiguousarray(gray, np.uint8)
        if g.ndim != 2:
            raise IamError("sift_detect wants a 2-D uint8 image")
        cap = int(max_out) if max_out > 0 else max(65536, g.size // 8)
        bufs = getattr(self, "_sift_out", None)
        if bufs is None or bufs[0].shape[0] < cap:       # output arrays are kept between calls (no fresh page faults)
            bufs = self._sift_out = (np.empty((cap, 5), np.float32), np.empty((cap,), np.int32), np.empty((cap, 128), np.uint8))
        kp, octv, des = bufs
        if max_out <= 0:
            cap = kp.shape[0]
        n = C.c_int(0)
        self._check(self._lib.iam_sift_detect(self._h, _ptr(g), g.shape[1], g.shape[0], cap, _ptr(kp), _ptr(octv), _ptr(des),
                                              C.byref(n)), "iam_sift_detect")
        return kp[:n.value].copy(), octv[:n.value].copy(), des[:n.value].copy()

    def debug_orb_fast(self, gray: np.ndarray) -> np.ndarray:
        g = np.ascontiguousarray(gray, np.uint8)
        out = np.zeros(g.shape, np.uint8)
        self._check(self._lib.iam_debug_orb_fast(self._h, _ptr(g), g.shape[1], g.shape[0], _ptr(out)), "iam_debug_orb_fast")
        return out
