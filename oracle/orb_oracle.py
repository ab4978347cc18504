"""ctypes bindings of the CPU oracle (liborb_oracle.so) — TEST INFRASTRUCTURE ONLY.

Stage functions mirror SURVEY.md §8.A0-A6 (cv::ORB as configured at
/root/reference/src/tracking/frame/stereo.rs:38-48) and §8.B rows a2-a8 (the reference's Rust loops).
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "liborb_oracle.so")

KP_DTYPE = np.dtype(
    [("x", "<f4"), ("y", "<f4"), ("size", "<f4"), ("angle", "<f4"), ("response", "<f4"), ("octave", "<i4"),
     ("class_id", "<i4")]
)
CORNER_DTYPE = np.dtype([("x", "<i4"), ("y", "<i4"), ("resp", "<f4")])
DMATCH_DTYPE = np.dtype([("query_idx", "<i4"), ("train_idx", "<i4"), ("img_idx", "<i4"), ("distance", "<f4")])
LEVEL_DTYPE = np.dtype([("w", "<i4"), ("h", "<i4"), ("scale", "<f4"), ("nfeatures", "<i4")])
assert KP_DTYPE.itemsize == 28 and CORNER_DTYPE.itemsize == 12 and DMATCH_DTYPE.itemsize == 16


class Camera(C.Structure):
    _fields_ = [("fx", C.c_double), ("fy", C.c_double), ("cx", C.c_double), ("cy", C.c_double),
                ("baseline", C.c_double)]


def build(force: bool = False) -> str:
    """Compile the oracle with the committed recipe (oracle/Makefile)."""
    srcs = [os.path.join(_HERE, f) for f in ("orb_oracle.cpp", "stereo_oracle.cpp", "classic_model.cpp", "brief_pattern.inc", "Makefile")]
    stale = (not os.path.exists(_LIB_PATH)) or any(os.path.getmtime(s) > os.path.getmtime(_LIB_PATH) for s in srcs)
    if force or stale:
        subprocess.run(["make", "-C", _HERE, "-B", "liborb_oracle.so"], check=True, capture_output=True)
    return _LIB_PATH


_lib = None


def lib():
    global _lib
    if _lib is None:
        build()
        _lib = C.CDLL(_LIB_PATH)
        _lib.orc_fast_atan2.restype = C.c_float
        _lib.orc_fast_atan2.argtypes = [C.c_float, C.c_float]
        _lib.orc_ic_angle.restype = C.c_float
        _lib.orc_descriptor_distance.restype = C.c_uint32
    return _lib


def _p(a):
    return a.ctypes.data_as(C.c_void_p)


def _u8(img):
    img = np.ascontiguousarray(img, dtype=np.uint8)
    assert img.ndim == 2
    return img


# ------------------------------------------------------------------ extraction stages (orb_oracle.cpp)
def level_geometry(W, H, nfeatures, nlevels=8, scale_factor=1.2):
    out = np.zeros(nlevels, dtype=LEVEL_DTYPE)
    lib().orc_level_geometry(int(W), int(H), int(nfeatures), int(nlevels), C.c_float(scale_factor), _p(out))
    return out


def resize_linear_exact(src, dw, dh):
    src = _u8(src)
    dst = np.empty((dh, dw), dtype=np.uint8)
    lib().orc_resize_linear_exact(_p(src), src.shape[1], src.shape[0], src.strides[0], _p(dst), dw, dh, dw)
    return dst


def build_pyramid(img, levels):
    pyr = [_u8(img)]
    for l in range(1, len(levels)):
        pyr.append(resize_linear_exact(pyr[-1], int(levels[l]["w"]), int(levels[l]["h"])))
    return pyr


def fast_nms(img, threshold=20):
    img = _u8(img)
    h, w = img.shape
    cap = w * h // 4 + 16
    out = np.zeros(cap, dtype=CORNER_DTYPE)
    n = lib().orc_fast_nms(_p(img), w, h, img.strides[0], int(threshold), _p(out), cap)
    assert n <= cap
    return out[:n].copy()


def border_filter(corners, w, h, border=31):
    v = corners.copy()
    n = lib().orc_border_filter(_p(v), len(v), int(w), int(h), int(border))
    return v[:n].copy()


def retain_best(corners, n_points):
    v = corners.copy()
    n = lib().orc_retain_best(_p(v), len(v), int(n_points))
    return v[:n].copy()


def harris(img, corners):
    img = _u8(img)
    v = corners.copy()
    lib().orc_harris(_p(img), img.strides[0], _p(v), len(v))
    return v


def ic_angles(img, corners):
    img = _u8(img)
    f = lib().orc_ic_angle
    return np.array([f(_p(img), img.strides[0], int(c["x"]), int(c["y"])) for c in corners], dtype=np.float32)


def fast_atan2(y, x):
    return float(lib().orc_fast_atan2(C.c_float(y), C.c_float(x)))


def blur7(img):
    img = _u8(img)
    dst = np.empty_like(img)
    lib().orc_blur7(_p(img), img.shape[1], img.shape[0], img.strides[0], _p(dst), dst.strides[0])
    return dst


def brief(blurred, corners, angles):
    blurred = _u8(blurred)
    desc = np.zeros((len(corners), 32), dtype=np.uint8)
    f = lib().orc_brief
    for i, c in enumerate(corners):
        f(_p(blurred), blurred.strides[0], int(c["x"]), int(c["y"]), C.c_float(float(angles[i])),
          C.c_void_p(desc.ctypes.data + 32 * i))
    return desc


def detect_and_compute(img, nfeatures=1200, nlevels=8, scale_factor=1.2, edge_threshold=31, fast_threshold=20):
    """The whole restated cv::ORB::detectAndCompute in one C++ call -> (keypoints[KP_DTYPE], desc[n,32])."""
    img = _u8(img)
    cap = 4 * nfeatures + 64
    kps = np.zeros(cap, dtype=KP_DTYPE)
    desc = np.zeros((cap, 32), dtype=np.uint8)
    n = lib().orc_detect_and_compute(_p(img), img.shape[1], img.shape[0], img.strides[0], int(nfeatures), int(nlevels),
                                     C.c_float(scale_factor), int(edge_threshold), int(fast_threshold), _p(kps),
                                     _p(desc), cap)
    assert n <= cap
    return kps[:n].copy(), desc[:n].copy()


def detect_stages(img, nfeatures=1200, nlevels=8, scale_factor=1.2, edge_threshold=31, fast_threshold=20):
    """Stage-by-stage run that keeps every intermediate (used to diff GPU kernels in isolation)."""
    img = _u8(img)
    H, W = img.shape
    lv = level_geometry(W, H, nfeatures, nlevels, scale_factor)
    pyr = build_pyramid(img, lv)
    st = dict(levels=lv, pyramid=pyr, fast=[], cand=[], cand_harris=[], final=[], angles=[], blurred=[], desc=[])
    kps, descs = [], []
    for l in range(nlevels):
        w, h = int(lv[l]["w"]), int(lv[l]["h"])
        c = border_filter(fast_nms(pyr[l], fast_threshold), w, h, edge_threshold)
        st["fast"].append(c)
        c = retain_best(c, 2 * int(lv[l]["nfeatures"]))
        st["cand"].append(c)
        c = harris(pyr[l], c)
        st["cand_harris"].append(c)
        c = retain_best(c, int(lv[l]["nfeatures"]))
        st["final"].append(c)
        ang = ic_angles(pyr[l], c)
        st["angles"].append(ang)
        b = blur7(pyr[l])
        st["blurred"].append(b)
        d = brief(b, c, ang)
        st["desc"].append(d)
        k = np.zeros(len(c), dtype=KP_DTYPE)
        s = np.float32(lv[l]["scale"])
        k["x"] = c["x"].astype(np.float32) * s
        k["y"] = c["y"].astype(np.float32) * s
        k["size"] = np.float32(31.0) * s
        k["angle"] = ang
        k["response"] = c["resp"]
        k["octave"] = l
        k["class_id"] = -1
        kps.append(k)
        descs.append(d)
    st["keypoints"] = np.concatenate(kps)
    st["descriptors"] = np.concatenate(descs) if descs else np.zeros((0, 32), np.uint8)
    return st


# ------------------------------------------------------------------ Rust-loop restatements (stereo_oracle.cpp)
def descriptor_distance(a, b):
    a = np.ascontiguousarray(a, dtype=np.uint8).ravel()
    b = np.ascontiguousarray(b, dtype=np.uint8).ravel()
    return int(lib().orc_descriptor_distance(_p(a), a.size, _p(b), b.size))


def _cam(camera):
    return Camera(**{k: float(camera[k]) for k in ("fx", "fy", "cx", "cy", "baseline")})


def match_features(kl, dl, kr, dr, camera):
    kl = np.ascontiguousarray(kl, dtype=KP_DTYPE)
    kr = np.ascontiguousarray(kr, dtype=KP_DTYPE)
    dl = np.ascontiguousarray(dl, dtype=np.uint8)
    dr = np.ascontiguousarray(dr, dtype=np.uint8)
    out = np.zeros(max(len(kl), 1), dtype=DMATCH_DTYPE)
    cam = _cam(camera)
    n = lib().orc_match_features(_p(kl), _p(dl), len(kl), _p(kr), _p(dr), len(kr), C.byref(cam), _p(out))
    return out[:n].copy()


def triangulate(kl, kr, matches, camera):
    kl = np.ascontiguousarray(kl, dtype=KP_DTYPE)
    kr = np.ascontiguousarray(kr, dtype=KP_DTYPE)
    m = np.ascontiguousarray(matches, dtype=DMATCH_DTYPE)
    pts = np.zeros((max(len(kl), 1), 3), dtype=np.float64)
    valid = np.zeros(max(len(kl), 1), dtype=np.uint8)
    cam = _cam(camera)
    lib().orc_triangulate(_p(kl), len(kl), _p(kr), len(kr), _p(m), len(m), C.byref(cam), _p(pts), _p(valid))
    return pts[: len(kl)], valid[: len(kl)]


def grid_build(kp, img_w=752.0, img_h=480.0):
    kp = np.ascontiguousarray(kp, dtype=KP_DTYPE)
    cs = np.zeros(64 * 48 + 1, dtype=np.int32)
    items = np.zeros(max(len(kp), 1), dtype=np.int32)
    lib().orc_grid_build(_p(kp), len(kp), C.c_double(img_w), C.c_double(img_h), _p(cs), _p(items))
    return cs, items[: len(kp)]


def grid_query(cs, items, x, y, r, img_w=752.0, img_h=480.0):
    out = np.zeros(max(len(items), 1), dtype=np.int32)
    items = np.ascontiguousarray(items, dtype=np.int32)
    n = lib().orc_grid_query(_p(cs), _p(items), C.c_double(x), C.c_double(y), C.c_double(r), C.c_double(img_w),
                             C.c_double(img_h), _p(out))
    return out[:n].copy()


def search_by_projection(kp, desc, uv, mp_desc, mode, radius=15.0, th_high=100, nn_ratio=0.75, img_w=752.0,
                         img_h=480.0, bound_w=0.0, bound_h=0.0):
    kp = np.ascontiguousarray(kp, dtype=KP_DTYPE)
    desc = np.ascontiguousarray(desc, dtype=np.uint8)
    uv = np.ascontiguousarray(uv, dtype=np.float64).reshape(-1, 2)
    mp_desc = np.ascontiguousarray(mp_desc, dtype=np.uint8).reshape(-1, 32)
    M = len(uv)
    idx = np.full(max(M, 1), -1, dtype=np.int32)
    dist = np.full(max(M, 1), -1, dtype=np.int32)
    lib().orc_search_by_projection(_p(kp), _p(desc), len(kp), C.c_double(img_w), C.c_double(img_h), _p(uv),
                                   _p(mp_desc), M, C.c_double(radius), C.c_uint32(th_high), C.c_float(nn_ratio),
                                   int(mode), C.c_double(bound_w), C.c_double(bound_h), _p(idx), _p(dist))
    return idx[:M], dist[:M]


def bf_match_crosscheck(query, train):
    q = np.ascontiguousarray(query, dtype=np.uint8).reshape(-1, 32)
    t = np.ascontiguousarray(train, dtype=np.uint8).reshape(-1, 32)
    out = np.zeros(max(len(q), 1), dtype=DMATCH_DTYPE)
    n = lib().orc_bf_match_crosscheck(_p(q), len(q), _p(t), len(t), _p(out))
    return out[:n].copy()


def fuse_search(kp, desc, uvr, mp_desc, desc_threshold=50):
    kp = np.ascontiguousarray(kp, dtype=KP_DTYPE)
    desc = np.ascontiguousarray(desc, dtype=np.uint8)
    uvr = np.ascontiguousarray(uvr, dtype=np.float64).reshape(-1, 3)
    mp_desc = np.ascontiguousarray(mp_desc, dtype=np.uint8).reshape(-1, 32)
    M = len(uvr)
    idx = np.full(max(M, 1), -1, dtype=np.int32)
    dist = np.full(max(M, 1), -1, dtype=np.int32)
    lib().orc_fuse_search(_p(kp), _p(desc), len(kp), _p(uvr), _p(mp_desc), M, C.c_uint32(desc_threshold), _p(idx), _p(dist))
    return idx[:M], dist[:M]


def distinctive_descriptors(desc, offsets):
    desc = np.ascontiguousarray(desc, dtype=np.uint8).reshape(-1, 32)
    offsets = np.ascontiguousarray(offsets, dtype=np.int32)
    M = len(offsets) - 1
    out = np.full(max(M, 1), -1, dtype=np.int32)
    lib().orc_distinctive_descriptors(_p(desc), _p(offsets), M, _p(out))
    return out[:M]


def make_vocabulary_arrays(parent_file_order, is_leaf, node_desc):
    """Restates the tree construction of OrbVocabulary::load_from_text (vocabulary/mod.rs:109-206) for nodes given in
    file order (node i+1 = line i): returns the flattened arrays of include/vslam_b200.h
    (node_desc[n_nodes,32], child_begin, child_ids, parent, word_id) with node 0 = root."""
    parent_file_order = np.asarray(parent_file_order, dtype=np.int64)
    n_nodes = len(parent_file_order) + 1
    desc = np.zeros((n_nodes, 32), dtype=np.uint8)
    desc[1:] = np.asarray(node_desc, dtype=np.uint8).reshape(-1, 32)
    parent = np.full(n_nodes, 0xFFFFFFFF, dtype=np.uint32)
    word = np.full(n_nodes, -1, dtype=np.int32)
    children = [[] for _ in range(n_nodes)]
    wc = 0
    for i, p in enumerate(parent_file_order):
        nid = i + 1
        parent[nid] = np.uint32(p)
        if is_leaf[i]:
            word[nid] = wc
            wc += 1
        if p < nid:  # "if (parent_id as usize) < nodes.len()" at the time the node is appended
            children[int(p)].append(nid)
    child_begin = np.zeros(n_nodes + 1, dtype=np.int32)
    child_begin[1:] = np.cumsum([len(c) for c in children])
    child_ids = np.array([c for cs in children for c in cs], dtype=np.int32)
    return desc, child_begin, child_ids, parent, word


def bow_quantize(desc, vocab, levels_up):
    """vocab = (node_desc, child_begin, child_ids, parent, word_id) -> (word, leaf, node_at_level) uint32 arrays."""
    nd, cb, ci, pa, wi = vocab
    desc = np.ascontiguousarray(desc, dtype=np.uint8).reshape(-1, 32)
    nd = np.ascontiguousarray(nd, dtype=np.uint8)
    cb = np.ascontiguousarray(cb, dtype=np.int32)
    ci = np.ascontiguousarray(ci if len(ci) else np.zeros(1), dtype=np.int32)
    pa = np.ascontiguousarray(pa, dtype=np.uint32)
    wi = np.ascontiguousarray(wi, dtype=np.int32)
    n = len(desc)
    w = np.zeros(max(n, 1), dtype=np.uint32)
    lf = np.zeros(max(n, 1), dtype=np.uint32)
    nl = np.zeros(max(n, 1), dtype=np.uint32)
    lib().orc_bow_quantize(_p(desc), n, _p(nd), _p(cb), _p(ci), _p(pa), _p(wi), int(levels_up), _p(w), _p(lf), _p(nl))
    return w[:n], lf[:n], nl[:n]


class TriParams(C.Structure):
    _fields_ = [("F", C.c_double * 9), ("epipole", C.c_double * 2), ("max_dist", C.c_uint32), ("radius", C.c_float),
                ("image_width", C.c_uint32), ("image_height", C.c_uint32)]


def tri_params(F, epipole, max_dist, camera, radius=100.0):
    """image_width/height as triangulation.rs:432-433: (camera.cx * 2.0) as u32."""
    p = TriParams()
    p.F[:] = [float(v) for v in np.asarray(F, dtype=np.float64).reshape(9)]
    p.epipole[:] = [float(epipole[0]), float(epipole[1])]
    p.max_dist, p.radius = int(max_dist), float(radius)
    p.image_width, p.image_height = int(camera["cx"] * 2.0), int(camera["cy"] * 2.0)
    return p


def _tri_args(kp1, desc1, flags1, kp2, desc2, flags2):
    return (np.ascontiguousarray(kp1, dtype=KP_DTYPE), np.ascontiguousarray(desc1, dtype=np.uint8).reshape(-1, 32),
            np.ascontiguousarray(flags1, dtype=np.uint8), np.ascontiguousarray(kp2, dtype=KP_DTYPE),
            np.ascontiguousarray(desc2, dtype=np.uint8).reshape(-1, 32), np.ascontiguousarray(flags2, dtype=np.uint8))


def search_for_triangulation(kp1, desc1, flags1, kp2, desc2, flags2, params):
    """triangulation.rs:401-530 -> int32 [n, 2] pairs (idx1, idx2) in emission order."""
    k1, d1, f1, k2, d2, f2 = _tri_args(kp1, desc1, flags1, kp2, desc2, flags2)
    out = np.zeros((max(len(k1), 1), 2), dtype=np.int32)
    n = lib().orc_search_for_triangulation(_p(k1), _p(d1), _p(f1), len(k1), _p(k2), _p(d2), _p(f2), len(k2),
                                           C.byref(params), _p(out))
    return out[:n].copy()


def csr(lists):
    start = np.zeros(len(lists) + 1, dtype=np.int32)
    start[1:] = np.cumsum([len(x) for x in lists])
    items = np.array([v for x in lists for v in x] or [0], dtype=np.int32)
    return start, items


def search_for_triangulation_bow(groups, kp1, desc1, flags1, kp2, desc2, flags2, params):
    """triangulation.rs:541-658; groups = [(indices1, indices2)] in the caller's HashMap visiting order."""
    k1, d1, f1, k2, d2, f2 = _tri_args(kp1, desc1, flags1, kp2, desc2, flags2)
    s1, i1 = csr([g[0] for g in groups])
    s2, i2 = csr([g[1] for g in groups])
    out = np.zeros((max(int(s1[-1]), 1), 2), dtype=np.int32)
    n = lib().orc_search_for_triangulation_bow(_p(k1), _p(d1), _p(f1), len(k1), _p(k2), _p(d2), _p(f2), len(k2),
                                               C.byref(params), len(groups), _p(s1), _p(i1), _p(s2), _p(i2), _p(out))
    return out[:n].copy()


# ------------------------------------------------------------------ classic mode model (classic_model.cpp) — NOT reference parity
def classic_detect_and_compute(img, nfeatures=1200, nlevels=8, scale_factor=1.2, ini_th=20, min_th=7):
    """Own CPU model of the upstream-ORB-SLAM3 style extraction (per-cell FAST with threshold retry + quadtree)."""
    img = _u8(img)
    cap = 4 * nfeatures + 256
    kps = np.zeros(cap, dtype=KP_DTYPE)
    desc = np.zeros((cap, 32), dtype=np.uint8)
    n = lib().orc_classic_detect_and_compute(_p(img), img.shape[1], img.shape[0], img.strides[0], int(nfeatures), int(nlevels),
                                             C.c_float(scale_factor), int(ini_th), int(min_th), _p(kps), _p(desc), cap)
    assert n <= cap
    return kps[:n].copy(), desc[:n].copy()


def classic_level_candidates(level_img, nfeat, ini_th=20, min_th=7):
    """-> (candidates int32 [n,3] (x, y, resp) in cell-row-major / raster order, kept indices) of one pyramid level."""
    img = _u8(level_img)
    cap = img.size // 4 + 16
    xyr = np.zeros((cap, 3), dtype=np.int32)
    keep = np.zeros(cap, dtype=np.int32)
    nk = C.c_int(0)
    n = lib().orc_classic_level_candidates(_p(img), img.shape[1], img.shape[0], img.strides[0], int(nfeat), int(ini_th),
                                           int(min_th), _p(xyr), cap, _p(keep), C.byref(nk))
    return xyr[:n].copy(), keep[: nk.value].copy()


def classic_stereo(kl, dl, kr, dr, left, right, camera, nfeatures=1200, nlevels=8, scale_factor=1.2):
    """Own CPU model of ComputeStereoMatches: row-band Hamming match (octave-scaled radius, +-1 level), 11x11 SAD over 11
    shifts on the keypoint's pyramid level, parabola, median filter -> (match_idx, orb_dist, u_right, depth) per left kp."""
    kl = np.ascontiguousarray(kl, dtype=KP_DTYPE)
    kr = np.ascontiguousarray(kr, dtype=KP_DTYPE)
    dl = np.ascontiguousarray(dl, dtype=np.uint8).reshape(-1, 32)
    dr = np.ascontiguousarray(dr, dtype=np.uint8).reshape(-1, 32)
    lv = level_geometry(left.shape[1], left.shape[0], nfeatures, nlevels, scale_factor)
    pl = np.concatenate([p.ravel() for p in build_pyramid(left, lv)])
    pr = np.concatenate([p.ravel() for p in build_pyramid(right, lv)])
    n = max(len(kl), 1)
    mi = np.full(n, -1, dtype=np.int32)
    od = np.full(n, -1, dtype=np.int32)
    ur = np.full(n, -1, dtype=np.float32)
    dp = np.full(n, -1, dtype=np.float32)
    lib().orc_classic_stereo(_p(kl), _p(dl), len(kl), _p(kr), _p(dr), len(kr), _p(pl), _p(pr), _p(lv), int(nlevels),
                             C.c_double(camera["fx"]), C.c_double(camera["baseline"]), _p(mi), _p(od), _p(ur), _p(dp))
    k = len(kl)
    return mi[:k], od[:k], ur[:k], dp[:k]
