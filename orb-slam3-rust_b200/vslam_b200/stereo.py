"""Host-side mirror of the reference's stereo front-end interface, over the C ABI (ctypes).

Names, argument meaning and error behaviour follow /root/reference/src/tracking/frame/{stereo,camera}.rs and
src/tracking/tracking_frame.rs so that tests read like tests of the reference:

    StereoProcessor(camera, n_features)            <- StereoProcessor::new           stereo.rs:37
    StereoProcessor.process(left, right, ts)       <- StereoProcessor::process       stereo.rs:52
    FeatureSet / StereoFrame / CameraModel         <- stereo.rs:15-29, camera.rs:3-10
    descriptor_distance(a, b)                      <- stereo.rs:166
    Frame.from_stereo(stereo_frame)                <- tracking_frame.rs:194
    FeatureGrid.get_features_in_area(x, y, r)      <- tracking_frame.rs:104
    search_by_projection(...)                      <- tracker.rs:863-923 / 1093-1158
    bf_match_crosscheck(query, train)              <- tracker.rs:1001-1010

Errors: the reference returns anyhow::Result; here a non-zero C status raises VslamError carrying
vslam_last_error().  There is no CPU path: without the CUDA library or a B200 every call raises.
"""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass, field

import numpy as np

from . import _abi
from ._abi import DMATCH_DTYPE, KP_DTYPE

TH_HIGH = 100  # stereo.rs:10
TH_LOW = 50  # stereo.rs:11
NN_RATIO = 0.75  # stereo.rs:12


class VslamError(RuntimeError):
    def __init__(self, status: int, message: str):
        super().__init__(f"vslam_b200 status {status}: {message}")
        self.status = status


@dataclass(frozen=True)
class CameraModel:  # camera.rs:3-10
    fx: float
    fy: float
    cx: float
    cy: float
    baseline: float

    @staticmethod
    def from_k_and_baseline(k, baseline):  # camera.rs:13-21
        k = np.asarray(k, dtype=np.float64)
        return CameraModel(float(k[0, 0]), float(k[1, 1]), float(k[0, 2]), float(k[1, 2]), float(baseline))


@dataclass
class FeatureSet:  # stereo.rs:15-19
    keypoints: np.ndarray  # KP_DTYPE [n]  (cv::KeyPoint layout)
    descriptors: np.ndarray  # uint8 [n, 32]


@dataclass
class StereoFrame:  # stereo.rs:21-29
    left_features: FeatureSet
    right_features: FeatureSet
    matches_lr: np.ndarray  # DMATCH_DTYPE [m], ascending query_idx
    points_cam: np.ndarray  # float64 [n_left, 3]
    points_valid: np.ndarray  # uint8 [n_left]  (Option tag of Vec<Option<Vector3<f64>>>)
    timestamp_ns: int = 0


def _ptr(a):
    return a.ctypes.data_as(C.c_void_p)


def _check_image(img, w, h, name):
    if not isinstance(img, np.ndarray) or img.dtype != np.uint8 or img.ndim != 2:
        raise VslamError(_abi.ERR_INVALID_ARGUMENT, f"{name}: expected a 2-D uint8 (8UC1) array")
    if img.shape != (h, w):
        raise VslamError(_abi.ERR_INVALID_ARGUMENT, f"{name}: expected {w}x{h}, got {img.shape[1]}x{img.shape[0]}")
    if img.strides[1] != 1:
        raise VslamError(_abi.ERR_INVALID_ARGUMENT, f"{name}: pixels of a row must be contiguous")
    return img


class _OutBuffers:
    """Caller-owned, capacity-checked output arrays for n frames (vslam_frame_out)."""

    def __init__(self, n, cap):
        self.n, self.cap = n, cap
        self.kpl = np.zeros((n, cap), dtype=KP_DTYPE)
        self.kpr = np.zeros((n, cap), dtype=KP_DTYPE)
        self.dl = np.zeros((n, cap, 32), dtype=np.uint8)
        self.dr = np.zeros((n, cap, 32), dtype=np.uint8)
        self.m = np.zeros((n, cap), dtype=DMATCH_DTYPE)
        self.pts = np.zeros((n, cap, 3), dtype=np.float64)
        self.valid = np.zeros((n, cap), dtype=np.uint8)
        self.outs = (_abi.FrameOut * n)()
        for i in range(n):
            o = self.outs[i]
            o.capacity = cap
            o.kp_left = self.kpl[i].ctypes.data
            o.kp_right = self.kpr[i].ctypes.data
            o.desc_left = self.dl[i].ctypes.data
            o.desc_right = self.dr[i].ctypes.data
            o.matches = self.m[i].ctypes.data
            o.points_cam = self.pts[i].ctypes.data
            o.points_valid = self.valid[i].ctypes.data

    def frame(self, i, copy=True):
        o = self.outs[i]
        nl, nr, nm = o.n_left, o.n_right, o.n_matches
        c = (lambda a: a.copy()) if copy else (lambda a: a)
        return StereoFrame(
            FeatureSet(c(self.kpl[i, :nl]), c(self.dl[i, :nl])),
            FeatureSet(c(self.kpr[i, :nr]), c(self.dr[i, :nr])),
            c(self.m[i, :nm]), c(self.pts[i, :nl]), c(self.valid[i, :nl]), int(o.timestamp_ns),
        )


class StereoProcessor:
    """B200 drop-in for the reference's StereoProcessor (stereo.rs:31-161)."""

    def __init__(self, camera: CameraModel, n_features: int, width: int = 752, height: int = 480, device: int = 0,
                 max_batch: int = 64, n_levels: int = 8, scale_factor: float = 1.2, edge_threshold: int = 31,
                 fast_threshold: int = 20, keypoint_capacity: int = 0, mode: str = "reference", classic_ini_fast: int = 0,
                 classic_min_fast: int = 0):
        self._lib = _abi.lib()
        self._h = C.c_void_p()
        cfg = _abi.Config()
        cfg.abi_version = _abi.ABI_VERSION
        cfg.width, cfg.height, cfg.n_features = int(width), int(height), int(n_features)
        cfg.n_levels, cfg.scale_factor = int(n_levels), float(scale_factor)
        cfg.edge_threshold, cfg.fast_threshold = int(edge_threshold), int(fast_threshold)
        cfg.camera = _abi.Camera(camera.fx, camera.fy, camera.cx, camera.cy, camera.baseline)
        cfg.device, cfg.max_batch, cfg.keypoint_capacity = int(device), int(max_batch), int(keypoint_capacity)
        if mode not in ("reference", "classic"):
            raise VslamError(_abi.ERR_INVALID_ARGUMENT, f"unknown mode {mode!r}")
        # "classic" = upstream-ORB-SLAM3 style per-cell FAST / quadtree / SAD + parabola: NOT reference parity (header)
        cfg.mode = _abi.MODE_CLASSIC if mode == "classic" else _abi.MODE_REFERENCE
        cfg.classic_ini_fast, cfg.classic_min_fast = int(classic_ini_fast), int(classic_min_fast)
        self.mode = mode
        rc = self._lib.vslam_create(C.byref(cfg), C.byref(self._h))
        if rc != _abi.OK:
            msg = self._lib.vslam_last_error(None).decode()
            self._h = C.c_void_p()
            raise VslamError(rc, msg)
        self.camera, self.n_features, self.width, self.height = camera, n_features, width, height
        self.max_batch = max_batch
        self.capacity = int(self._lib.vslam_keypoint_capacity(self._h))
        self.n_levels = n_levels

    # -- lifetime
    def close(self):
        if getattr(self, "_h", None) is not None and self._h.value:
            self._lib.vslam_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()

    def _raise(self, rc):
        raise VslamError(rc, self._lib.vslam_last_error(self._h).decode())

    # -- StereoProcessor::process (stereo.rs:52-66)
    def process(self, left: np.ndarray, right: np.ndarray, timestamp_ns: int = 0) -> StereoFrame:
        left = _check_image(left, self.width, self.height, "left")
        right = _check_image(right, self.width, self.height, "right")
        ob = _OutBuffers(1, self.capacity)
        rc = self._lib.vslam_process(self._h, _ptr(left), left.strides[0], _ptr(right), right.strides[0],
                                     int(timestamp_ns), ob.outs)
        if rc != _abi.OK:
            self._raise(rc)
        return ob.frame(0)

    # -- offline batch form of the caller loop (main.rs:66-85)
    def process_batch(self, lefts, rights, timestamps_ns=None, out: _OutBuffers | None = None, copy=True):
        """lefts/rights: uint8 arrays [n, H, W] (or lists of [H, W]) in HOST memory -> list[StereoFrame]."""
        n = len(lefts)
        if n != len(rights):
            raise VslamError(_abi.ERR_INVALID_ARGUMENT, "left/right batch sizes differ")
        if n == 0:
            return []
        lp = (C.c_void_p * n)()
        rp = (C.c_void_p * n)()
        stride = None
        for i in range(n):
            l = _check_image(lefts[i], self.width, self.height, f"left[{i}]")
            r = _check_image(rights[i], self.width, self.height, f"right[{i}]")
            if stride is None:
                stride = l.strides[0]
            if l.strides[0] != stride or r.strides[0] != stride:
                raise VslamError(_abi.ERR_INVALID_ARGUMENT, "all images of a batch must share one row stride")
            lp[i], rp[i] = l.ctypes.data, r.ctypes.data
        ts = None
        if timestamps_ns is not None:
            ts = np.ascontiguousarray(timestamps_ns, dtype=np.uint64)
        ob = out or _OutBuffers(n, self.capacity)
        rc = self._lib.vslam_process_batch(self._h, n, lp, rp, stride, _ptr(ts) if ts is not None else None, ob.outs)
        if rc != _abi.OK:
            self._raise(rc)
        return [ob.frame(i, copy) for i in range(n)]

    def new_output_buffers(self, n):
        return _OutBuffers(n, self.capacity)

    def process_batch_raw(self, n, left_ptrs, right_ptrs, stride, out: _OutBuffers):
        """Zero-overhead variant for benchmarking: pre-built pointer arrays, results land in `out`."""
        rc = self._lib.vslam_process_batch(self._h, n, left_ptrs, right_ptrs, stride, None, out.outs)
        if rc != _abi.OK:
            self._raise(rc)

    def process_batch_device(self, n, left_dev_ptr, right_dev_ptr, image_stride, row_stride, out: _OutBuffers | None):
        """Images already resident in HBM (device pointers as ints).  out=None keeps results on the device."""
        rc = self._lib.vslam_process_batch_device(self._h, int(n), C.c_void_p(left_dev_ptr), C.c_void_p(right_dev_ptr),
                                                  int(image_stride), int(row_stride), None,
                                                  out.outs if out is not None else None)
        if rc != _abi.OK:
            self._raise(rc)

    # -- detect_features (stereo.rs:68-78)
    def detect_features(self, image: np.ndarray) -> FeatureSet:
        image = _check_image(image, self.width, self.height, "image")
        kps = np.zeros(self.capacity, dtype=KP_DTYPE)
        desc = np.zeros((self.capacity, 32), dtype=np.uint8)
        n = C.c_int(0)
        rc = self._lib.vslam_detect_and_compute(self._h, _ptr(image), image.strides[0], _ptr(kps), _ptr(desc),
                                                self.capacity, C.byref(n))
        if rc != _abi.OK:
            self._raise(rc)
        return FeatureSet(kps[: n.value].copy(), desc[: n.value].copy())

    # -- match_features (stereo.rs:80-161)
    def match_features(self, left: FeatureSet, right: FeatureSet) -> np.ndarray:
        kl = np.ascontiguousarray(left.keypoints, dtype=KP_DTYPE)
        kr = np.ascontiguousarray(right.keypoints, dtype=KP_DTYPE)
        dl = np.ascontiguousarray(left.descriptors, dtype=np.uint8)
        dr = np.ascontiguousarray(right.descriptors, dtype=np.uint8)
        m = np.zeros(max(len(kl), 1), dtype=DMATCH_DTYPE)
        n = C.c_int(0)
        rc = self._lib.vslam_match_features(self._h, _ptr(kl), _ptr(dl), len(kl), _ptr(kr), _ptr(dr), len(kr), _ptr(m),
                                            C.byref(n))
        if rc != _abi.OK:
            self._raise(rc)
        return m[: n.value].copy()

    # -- triangulate (stereo.rs:186-216)
    def triangulate(self, left: FeatureSet, right: FeatureSet, matches: np.ndarray):
        kl = np.ascontiguousarray(left.keypoints, dtype=KP_DTYPE)
        kr = np.ascontiguousarray(right.keypoints, dtype=KP_DTYPE)
        m = np.ascontiguousarray(matches, dtype=DMATCH_DTYPE)
        pts = np.zeros((max(len(kl), 1), 3), dtype=np.float64)
        valid = np.zeros(max(len(kl), 1), dtype=np.uint8)
        rc = self._lib.vslam_triangulate(self._h, _ptr(kl), len(kl), _ptr(kr), len(kr), _ptr(m), len(m), _ptr(pts),
                                         _ptr(valid))
        if rc != _abi.OK:
            self._raise(rc)
        return pts[: len(kl)], valid[: len(kl)]

    # -- descriptor_distance (stereo.rs:166-175), batched
    def descriptor_distances(self, a: np.ndarray, b: np.ndarray) -> np.ndarray:
        a = np.ascontiguousarray(a, dtype=np.uint8).reshape(-1, 32)
        b = np.ascontiguousarray(b, dtype=np.uint8).reshape(-1, 32)
        if len(a) != len(b):
            raise VslamError(_abi.ERR_INVALID_ARGUMENT, "descriptor batches differ in length")
        out = np.zeros(max(len(a), 1), dtype=np.uint32)
        rc = self._lib.vslam_descriptor_distance(self._h, _ptr(a), _ptr(b), len(a), _ptr(out))
        if rc != _abi.OK:
            self._raise(rc)
        return out[: len(a)]

    # -- FeatureGrid (tracking_frame.rs:52-128)
    def grid_build(self, keypoints, img_w=752.0, img_h=480.0):
        kp = np.ascontiguousarray(keypoints, dtype=KP_DTYPE)
        cs = np.zeros(_abi.GRID_COLS * _abi.GRID_ROWS + 1, dtype=np.int32)
        items = np.zeros(max(len(kp), 1), dtype=np.int32)
        rc = self._lib.vslam_grid_build(self._h, _ptr(kp), len(kp), float(img_w), float(img_h), _ptr(cs), _ptr(items))
        if rc != _abi.OK:
            self._raise(rc)
        return cs, items[: len(kp)]

    def grid_query(self, keypoints, xyr, img_w=752.0, img_h=480.0):
        kp = np.ascontiguousarray(keypoints, dtype=KP_DTYPE)
        q = np.ascontiguousarray(xyr, dtype=np.float64).reshape(-1, 3)
        cap = max(len(kp) * max(len(q), 1), 1)
        start = np.zeros(len(q) + 1, dtype=np.int32)
        items = np.zeros(cap, dtype=np.int32)
        rc = self._lib.vslam_grid_query(self._h, _ptr(kp), len(kp), float(img_w), float(img_h), _ptr(q), len(q),
                                        _ptr(start), _ptr(items), cap)
        if rc != _abi.OK:
            self._raise(rc)
        return [items[start[i]: start[i + 1]].copy() for i in range(len(q))]

    # -- projection searches (tracker.rs:863-923, 1093-1158)
    def search_by_projection(self, features: FeatureSet, uv, mp_desc, mode=_abi.SEARCH_LOCAL_MAP, radius=15.0,
                             th_high=TH_HIGH, nn_ratio=NN_RATIO, img_w=752.0, img_h=480.0, bound_w=None, bound_h=None):
        kp = np.ascontiguousarray(features.keypoints, dtype=KP_DTYPE)
        d = np.ascontiguousarray(features.descriptors, dtype=np.uint8)
        uv = np.ascontiguousarray(uv, dtype=np.float64).reshape(-1, 2)
        md = np.ascontiguousarray(mp_desc, dtype=np.uint8).reshape(-1, 32)
        M = len(uv)
        sp = _abi.SearchParams(int(mode), float(radius), int(th_high), float(nn_ratio), float(img_w), float(img_h),
                               float(2.0 * self.camera.cx if bound_w is None else bound_w),
                               float(2.0 * self.camera.cy if bound_h is None else bound_h))
        idx = np.full(max(M, 1), -1, dtype=np.int32)
        dist = np.full(max(M, 1), -1, dtype=np.int32)
        rc = self._lib.vslam_search_by_projection(self._h, _ptr(kp), _ptr(d), len(kp), _ptr(uv), _ptr(md), M,
                                                  C.byref(sp), _ptr(idx), _ptr(dist))
        if rc != _abi.OK:
            self._raise(rc)
        return idx[:M], dist[:M]

    # -- frames resident in HBM: Frame::from_stereo (tracking_frame.rs:194-215) without a PCIe round trip of the features
    def frame_retain(self, index=0, img_w=752.0, img_h=480.0) -> int:
        """Keeps the left features + FeatureGrid of frame `index` of the last process* call on the device -> frame id."""
        fid = C.c_int(-1)
        rc = self._lib.vslam_frame_retain(self._h, int(index), float(img_w), float(img_h), C.byref(fid))
        if rc != _abi.OK:
            self._raise(rc)
        return fid.value

    def frame_upload(self, features: FeatureSet, img_w=752.0, img_h=480.0) -> int:
        kp = np.ascontiguousarray(features.keypoints, dtype=KP_DTYPE)
        d = np.ascontiguousarray(features.descriptors, dtype=np.uint8).reshape(-1, 32)
        fid = C.c_int(-1)
        rc = self._lib.vslam_frame_upload(self._h, _ptr(kp), _ptr(d), len(kp), float(img_w), float(img_h), C.byref(fid))
        if rc != _abi.OK:
            self._raise(rc)
        return fid.value

    def frame_release(self, frame_id):
        rc = self._lib.vslam_frame_release(self._h, int(frame_id))
        if rc != _abi.OK:
            self._raise(rc)

    def frame_num_features(self, frame_id):
        return int(self._lib.vslam_frame_num_features(self._h, int(frame_id)))

    def frame_search_by_projection(self, frame_id, uv, mp_desc, mode=_abi.SEARCH_LOCAL_MAP, radius=15.0, th_high=TH_HIGH,
                                   nn_ratio=NN_RATIO, img_w=752.0, img_h=480.0, bound_w=None, bound_h=None):
        """tracker.rs:863-923 / 1093-1158 against a resident frame: only (u, v) + map-point descriptors go up."""
        uv = np.ascontiguousarray(uv, dtype=np.float64).reshape(-1, 2)
        md = np.ascontiguousarray(mp_desc, dtype=np.uint8).reshape(-1, 32)
        M = len(uv)
        sp = _abi.SearchParams(int(mode), float(radius), int(th_high), float(nn_ratio), float(img_w), float(img_h),
                               float(2.0 * self.camera.cx if bound_w is None else bound_w),
                               float(2.0 * self.camera.cy if bound_h is None else bound_h))
        idx = np.full(max(M, 1), -1, dtype=np.int32)
        dist = np.full(max(M, 1), -1, dtype=np.int32)
        rc = self._lib.vslam_frame_search_by_projection(self._h, int(frame_id), _ptr(uv), _ptr(md), M, C.byref(sp),
                                                        _ptr(idx), _ptr(dist))
        if rc != _abi.OK:
            self._raise(rc)
        return idx[:M], dist[:M]

    def frame_grid_query(self, frame_id, xyr):
        q = np.ascontiguousarray(xyr, dtype=np.float64).reshape(-1, 3)
        n = max(self.frame_num_features(frame_id), 0)
        cap = max(n * max(len(q), 1), 1)
        start = np.zeros(len(q) + 1, dtype=np.int32)
        items = np.zeros(cap, dtype=np.int32)
        rc = self._lib.vslam_frame_grid_query(self._h, int(frame_id), _ptr(q), len(q), _ptr(start), _ptr(items), cap)
        if rc != _abi.OK:
            self._raise(rc)
        return [items[start[i]: start[i + 1]].copy() for i in range(len(q))]

    def frame_bf_match_crosscheck(self, query_frame_id, train_frame_id):
        nq = max(self.frame_num_features(query_frame_id), 1)
        m = np.zeros(nq, dtype=DMATCH_DTYPE)
        n = C.c_int(0)
        rc = self._lib.vslam_frame_bf_match_crosscheck(self._h, int(query_frame_id), int(train_frame_id), _ptr(m), C.byref(n))
        if rc != _abi.OK:
            self._raise(rc)
        return m[: n.value].copy()

    def last_search_ms(self):
        return float(self._lib.vslam_last_search_ms(self._h))

    # -- BFMatcher(NORM_HAMMING, crossCheck=true) (tracker.rs:1001-1010)
    def bf_match_crosscheck(self, query, train):
        q = np.ascontiguousarray(query, dtype=np.uint8).reshape(-1, 32)
        t = np.ascontiguousarray(train, dtype=np.uint8).reshape(-1, 32)
        m = np.zeros(max(len(q), 1), dtype=DMATCH_DTYPE)
        n = C.c_int(0)
        rc = self._lib.vslam_bf_match_crosscheck(self._h, _ptr(q), len(q), _ptr(t), len(t), _ptr(m), C.byref(n))
        if rc != _abi.OK:
            self._raise(rc)
        return m[: n.value].copy()

    # -- keyframe-rate consumers
    def fuse_search(self, features: FeatureSet, uvr, mp_desc, desc_threshold=TH_LOW):
        """Descriptor search of fuse_points_into_keyframes (search_in_neighbors.rs:303-340): uvr = (u, v, radius)."""
        kp = np.ascontiguousarray(features.keypoints, dtype=KP_DTYPE)
        d = np.ascontiguousarray(features.descriptors, dtype=np.uint8)
        q = np.ascontiguousarray(uvr, dtype=np.float64).reshape(-1, 3)
        md = np.ascontiguousarray(mp_desc, dtype=np.uint8).reshape(-1, 32)
        M = len(q)
        idx = np.full(max(M, 1), -1, dtype=np.int32)
        dist = np.full(max(M, 1), -1, dtype=np.int32)
        rc = self._lib.vslam_fuse_search(self._h, _ptr(kp), _ptr(d), len(kp), _ptr(q), _ptr(md), M, int(desc_threshold),
                                         _ptr(idx), _ptr(dist))
        if rc != _abi.OK:
            self._raise(rc)
        return idx[:M], dist[:M]

    def distinctive_descriptors(self, desc, offsets):
        """Map::compute_distinctive_descriptors (map.rs:917-937) for many map points at once (CSR groups)."""
        d = np.ascontiguousarray(desc, dtype=np.uint8).reshape(-1, 32)
        off = np.ascontiguousarray(offsets, dtype=np.int32)
        M = len(off) - 1
        out = np.full(max(M, 1), -1, dtype=np.int32)
        rc = self._lib.vslam_distinctive_descriptors(self._h, _ptr(d), _ptr(off), M, _ptr(out))
        if rc != _abi.OK:
            self._raise(rc)
        return out[:M]

    def vocab_upload(self, node_desc, child_begin, child_ids, parent, word_id):
        """Keeps a flattened OrbVocabulary tree (vocabulary/mod.rs:109-206) on the device; node 0 is the root."""
        nd = np.ascontiguousarray(node_desc, dtype=np.uint8).reshape(-1, 32)
        cb = np.ascontiguousarray(child_begin, dtype=np.int32)
        ci = np.ascontiguousarray(child_ids if len(child_ids) else np.zeros(1), dtype=np.int32)
        pa = np.ascontiguousarray(parent, dtype=np.uint32)
        wi = np.ascontiguousarray(word_id, dtype=np.int32)
        if not (len(cb) == len(nd) + 1 and len(pa) == len(nd) and len(wi) == len(nd)):
            raise ValueError("vocabulary arrays disagree on the node count")
        rc = self._lib.vslam_vocab_upload(self._h, len(nd), _ptr(nd), _ptr(cb), _ptr(ci), _ptr(pa), _ptr(wi))
        if rc != _abi.OK:
            self._raise(rc)

    def bow_quantize(self, desc, levels_up):
        """OrbVocabulary::transform_descriptor + get_parent_at_level (vocabulary/mod.rs:230-275) for every row of desc:
        returns (word_id, leaf_node, node_at_level) as uint32 arrays."""
        d = np.ascontiguousarray(desc, dtype=np.uint8).reshape(-1, 32)
        n = len(d)
        w = np.zeros(max(n, 1), dtype=np.uint32)
        lf = np.zeros(max(n, 1), dtype=np.uint32)
        nl = np.zeros(max(n, 1), dtype=np.uint32)
        rc = self._lib.vslam_bow_quantize(self._h, _ptr(d), n, int(levels_up), _ptr(w), _ptr(lf), _ptr(nl))
        if rc != _abi.OK:
            self._raise(rc)
        return w[:n], lf[:n], nl[:n]

    def search_for_triangulation(self, kf1: FeatureSet, flags1, kf2: FeatureSet, flags2, F, epipole, max_dist,
                                 groups=None, radius=100.0):
        """search_for_triangulation (local_mapping/triangulation.rs:401-530); with `groups` = [(indices1, indices2)]
        in the caller's FeatureVector visiting order: search_for_triangulation_bow (:541-658).  F and the epipole are
        computed by the caller (:417-428, :668-680).  flags: bit 0 = has a map point, bit 1 (kf1) = has stereo depth.
        -> int32 [n, 2] pairs (idx1, idx2) in emission order."""
        k1 = np.ascontiguousarray(kf1.keypoints, dtype=KP_DTYPE)
        d1 = np.ascontiguousarray(kf1.descriptors, dtype=np.uint8).reshape(-1, 32)
        f1 = np.ascontiguousarray(flags1, dtype=np.uint8)
        k2 = np.ascontiguousarray(kf2.keypoints, dtype=KP_DTYPE)
        d2 = np.ascontiguousarray(kf2.descriptors, dtype=np.uint8).reshape(-1, 32)
        f2 = np.ascontiguousarray(flags2, dtype=np.uint8)
        tp = _abi.TriParams()
        tp.F[:] = [float(v) for v in np.asarray(F, dtype=np.float64).reshape(9)]
        tp.epipole[:] = [float(epipole[0]), float(epipole[1])]
        tp.max_dist, tp.radius = int(max_dist), float(radius)
        tp.image_width, tp.image_height = int(self.camera.cx * 2.0), int(self.camera.cy * 2.0)  # `as u32`
        n = C.c_int(0)
        if groups is None:
            out = np.zeros((max(len(k1), 1), 2), dtype=np.int32)
            rc = self._lib.vslam_search_for_triangulation(self._h, _ptr(k1), _ptr(d1), _ptr(f1), len(k1), _ptr(k2),
                                                          _ptr(d2), _ptr(f2), len(k2), C.byref(tp), _ptr(out), C.byref(n))
        else:
            def csr(lists):
                start = np.zeros(len(lists) + 1, dtype=np.int32)
                start[1:] = np.cumsum([len(x) for x in lists])
                return start, np.array([v for x in lists for v in x] or [0], dtype=np.int32)

            s1, i1 = csr([g[0] for g in groups])
            s2, i2 = csr([g[1] for g in groups])
            out = np.zeros((max(int(s1[-1]), 1), 2), dtype=np.int32)
            rc = self._lib.vslam_search_for_triangulation_bow(self._h, _ptr(k1), _ptr(d1), _ptr(f1), len(k1), _ptr(k2),
                                                              _ptr(d2), _ptr(f2), len(k2), C.byref(tp), len(groups),
                                                              _ptr(s1), _ptr(i1), _ptr(s2), _ptr(i2), _ptr(out),
                                                              C.byref(n))
        if rc != _abi.OK:
            self._raise(rc)
        return out[: n.value].copy()

    # -- classic mode extras (NOT reference parity; see include/vslam_b200.h)
    def classic_stereo_extra(self, index=0):
        """(u_right, depth) float32 per left keypoint of frame `index` of the last process* call; -1 = no stereo match."""
        ur = np.full(self.capacity, -1, dtype=np.float32)
        dp = np.full(self.capacity, -1, dtype=np.float32)
        rc = self._lib.vslam_classic_stereo_extra(self._h, int(index), _ptr(ur), _ptr(dp), self.capacity)
        if rc != _abi.OK:
            self._raise(rc)
        return ur, dp

    def debug_classic_candidates(self, image, level):
        w, h, _, _ = self.level_info(level)
        cap = w * h // 4 + 1024
        buf = np.zeros((cap, 3), dtype=np.int32)
        n = C.c_int(0)
        rc = self._lib.vslam_debug_classic_candidates(self._h, image, level, _ptr(buf), cap, C.byref(n))
        if rc != _abi.OK:
            self._raise(rc)
        return buf[: n.value].copy()

    # -- introspection (parity tests / profiling)
    def level_info(self, level):
        w, h, s, n = C.c_int(), C.c_int(), C.c_float(), C.c_int()
        rc = self._lib.vslam_level_info(self._h, level, C.byref(w), C.byref(h), C.byref(s), C.byref(n))
        if rc != _abi.OK:
            self._raise(rc)
        return w.value, h.value, s.value, n.value

    def debug_level_image(self, stage, image, level):
        w, h, _, _ = self.level_info(level)
        dst = np.zeros((h, w), dtype=np.uint8)
        rc = self._lib.vslam_debug_level_image(self._h, stage, image, level, _ptr(dst), dst.size)
        if rc != _abi.OK:
            self._raise(rc)
        return dst

    def debug_fast_corners(self, image, level):
        w, h, _, _ = self.level_info(level)
        cap = w * h // 4 + 16
        buf = np.zeros((cap, 3), dtype=np.int32)
        n = C.c_int(0)
        rc = self._lib.vslam_debug_fast_corners(self._h, image, level, _ptr(buf), cap, C.byref(n))
        if rc != _abi.OK:
            self._raise(rc)
        return buf[: n.value].copy()

    def kernel_launches(self):
        return int(self._lib.vslam_kernel_launches(self._h))

    def set_profiling(self, enabled):
        self._lib.vslam_set_profiling(self._h, int(bool(enabled)))

    def last_batch_ms(self):
        """Device time (CUDA events on the library's streams) of the last process_batch* call."""
        return float(self._lib.vslam_last_batch_ms(self._h))

    def last_stage_ms(self):
        ms = np.zeros(7, dtype=np.float32)
        self._lib.vslam_last_stage_ms(self._h, _ptr(ms))
        return ms


def descriptor_distance(processor: StereoProcessor, desc1, desc2) -> int:
    """descriptor_distance(&desc1, &desc2) (stereo.rs:166): Hamming distance of two 32-byte descriptors."""
    return int(processor.descriptor_distances(np.asarray(desc1).reshape(1, 32), np.asarray(desc2).reshape(1, 32))[0])


class FeatureGrid:
    """FeatureGrid (tracking_frame.rs:23-128): 64x48 cells; built and queried on the GPU."""

    GRID_COLS, GRID_ROWS = 64, 48

    def __init__(self, processor: StereoProcessor, keypoints, img_width=752.0, img_height=480.0):
        self._p, self._kp, self._w, self._h = processor, keypoints, float(img_width), float(img_height)
        self.cell_start, self.cell_items = processor.grid_build(keypoints, img_width, img_height)

    def get_features_in_area(self, x, y, r):
        return self._p.grid_query(self._kp, [[x, y, r]], self._w, self._h)[0]


@dataclass
class Frame:
    """Tracking Frame (tracking_frame.rs:172-226)."""

    timestamp_ns: int
    features: FeatureSet
    points_cam: np.ndarray
    points_valid: np.ndarray
    grid: FeatureGrid
    map_point_matches: list = field(default_factory=list)

    @staticmethod
    def from_stereo(processor: StereoProcessor, stereo: StereoFrame):
        # EuRoC dimensions are hard-coded in the reference (tracking_frame.rs:194-197)
        return Frame.from_stereo_with_size(processor, stereo, 752.0, 480.0)

    @staticmethod
    def from_stereo_with_size(processor, stereo, img_width, img_height):
        n = len(stereo.left_features.keypoints)
        grid = FeatureGrid(processor, stereo.left_features.keypoints, img_width, img_height)
        return Frame(stereo.timestamp_ns, stereo.left_features, stereo.points_cam, stereo.points_valid, grid, [None] * n)

    def num_features(self):
        return len(self.features.keypoints)

    def keypoints(self):
        return self.features.keypoints
