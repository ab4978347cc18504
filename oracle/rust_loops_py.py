"""Second, INDEPENDENT restatement of the reference's Rust loops in pure Python — TEST INFRASTRUCTURE ONLY.

Written straight from the Rust sources, without looking at oracle/stereo_oracle.cpp, so that a transcription error in one
restatement does not silently pass: tests/test_oracle_differential.py fuzzes the two against each other (the Rust
crate cannot be built in this image, so this is the only independent check available for rows a2-a4, a7, a8, f2).

Rust semantics coded explicitly:
  * `as usize` / `as i32` / `as u32` from a float SATURATES (negative -> 0 for unsigned, NaN -> 0);
  * `as usize` from a negative i32 WRAPS (sign extension to 64 bits);
  * f32 arithmetic is done on numpy float32 scalars (one IEEE rounding per operation, no fused multiply-add);
  * integer compare chains are strict `<` in scan order (first index attaining the minimum wins).

Every function cites the reference lines it follows (paths relative to /root/reference/).
"""
from __future__ import annotations

import math

import numpy as np

F32 = np.float32
TH_HIGH = 100  # src/tracking/frame/stereo.rs:10
TH_LOW = 50  # stereo.rs:11
NN_RATIO = F32(0.75)  # stereo.rs:12
U32_MAX = 0xFFFFFFFF
USIZE_BITS = 64


def f64_as_usize(v: float) -> int:
    if v != v or v <= 0.0:
        return 0
    if v >= 2.0 ** USIZE_BITS:
        return 2 ** USIZE_BITS - 1
    return int(v)  # truncation toward zero


def f64_as_i32(v: float) -> int:
    if v != v:
        return 0
    if v <= -(2.0 ** 31):
        return -(2 ** 31)
    if v >= 2.0 ** 31 - 1:
        return 2 ** 31 - 1
    return int(v)


def f64_as_u32(v: float) -> int:
    if v != v or v <= 0.0:
        return 0
    if v >= 2.0 ** 32 - 1:
        return 2 ** 32 - 1
    return int(v)


def i32_as_usize(v: int) -> int:
    return v if v >= 0 else v + 2 ** USIZE_BITS


# ------------------------------------------------------------------------------------------------- stereo.rs
def descriptor_distance(d1, d2) -> int:
    """stereo.rs:166-175: sum over min(cols) bytes of popcount(a ^ b)."""
    cols = min(len(d1), len(d2))
    total = 0
    for j in range(cols):
        total += bin(int(d1[j]) ^ int(d2[j])).count("1")
    return total


def match_features(kp_l, desc_l, kp_r, desc_r, cam):
    """stereo.rs:80-161 -> list of (query_idx, train_idx, img_idx, distance)."""
    max_disparity = F32(cam["fx"] * cam["baseline"] / 0.1)  # f64 expression, then `as f32`
    min_disparity = F32(cam["fx"] * cam["baseline"] / 40.0)
    n_l, n_r = len(kp_l), len(kp_r)
    out = []
    for left_idx in range(n_l):
        ul, vl = F32(kp_l["x"][left_idx]), F32(kp_l["y"][left_idx])
        min_u = max(F32(ul - max_disparity), F32(0.0))
        cap = F32(F32(F32(n_r) * ul) / F32(n_l))  # (len as f32) * ul / (len as f32)
        max_u = min(F32(ul - min_disparity), cap)
        best_dist, best_right, second = TH_HIGH, None, TH_HIGH
        for right_idx in range(n_r):
            ur, vr = F32(kp_r["x"][right_idx]), F32(kp_r["y"][right_idx])
            if abs(F32(vl - vr)) > F32(2.0):
                continue
            if ur < min_u or ur > max_u:
                continue
            if ul <= ur:
                continue
            dist = descriptor_distance(desc_l[left_idx], desc_r[right_idx])
            if dist < best_dist:
                second, best_dist, best_right = best_dist, dist, right_idx
            elif dist < second:
                second = dist
        if best_right is not None:
            if F32(best_dist) < F32(F32(0.9) * F32(second)) or second == TH_HIGH:
                out.append((left_idx, best_right, 0, float(best_dist)))
    return out


def triangulate(kp_l, kp_r, matches, cam):
    """stereo.rs:186-216 -> list of None | (x, y, z) per left keypoint (f64)."""
    pts = [None] * len(kp_l)
    for (q, t, _img, _d) in matches:
        if not (0 <= q < len(kp_l) and 0 <= t < len(kp_r)):
            continue  # .get() -> None
        lx, ly = float(kp_l["x"][q]), float(kp_l["y"][q])
        rx = float(kp_r["x"][t])
        disparity = lx - rx
        if abs(disparity) < 0.5:
            continue
        z = cam["fx"] * cam["baseline"] / disparity
        x = (lx - cam["cx"]) * z / cam["fx"]
        y = (ly - cam["cy"]) * z / cam["fy"]
        pts[q] = (x, y, z)
    return pts


# ----------------------------------------------------------------------------------------- tracking_frame.rs
GRID_COLS, GRID_ROWS = 64, 48  # tracking_frame.rs:43-44


class FeatureGrid:
    """tracking_frame.rs:52-128."""

    def __init__(self, kp, img_width=752.0, img_height=480.0):
        self.min_x = self.min_y = 0.0
        self.w_inv = GRID_COLS / (img_width - self.min_x)
        self.h_inv = GRID_ROWS / (img_height - self.min_y)
        self.cells = [[] for _ in range(GRID_COLS * GRID_ROWS)]
        for idx in range(len(kp)):
            x, y = float(kp["x"][idx]), float(kp["y"][idx])
            cx = min(f64_as_usize((x - self.min_x) * self.w_inv), GRID_COLS - 1)
            cy = min(f64_as_usize((y - self.min_y) * self.h_inv), GRID_ROWS - 1)
            self.cells[cy * GRID_COLS + cx].append(idx)

    def get_features_in_area(self, x, y, r):
        min_cx = f64_as_i32(math.floor((x - self.min_x - r) * self.w_inv))
        max_cx = f64_as_i32(math.ceil((x - self.min_x + r) * self.w_inv))
        min_cy = f64_as_i32(math.floor((y - self.min_y - r) * self.h_inv))
        max_cy = f64_as_i32(math.ceil((y - self.min_y + r) * self.h_inv))
        min_cx = i32_as_usize(max(min_cx, 0))
        max_cx = min(i32_as_usize(max_cx), GRID_COLS - 1)  # the cast happens BEFORE the clamp: negatives wrap
        min_cy = i32_as_usize(max(min_cy, 0))
        max_cy = min(i32_as_usize(max_cy), GRID_ROWS - 1)
        out = []
        for cy in range(min_cy, max_cy + 1):
            for cx in range(min_cx, max_cx + 1):
                out.extend(self.cells[cy * GRID_COLS + cx])
        return out


# ------------------------------------------------------------------------------------------------ tracker.rs
def search_local_map(grid, desc, uv, mp_desc, radius=15.0):
    """tracker.rs:863-923 (per map point; the projection and the z > 0 test stay with the caller) -> [kp_idx | None]."""
    out = []
    for m in range(len(uv)):
        cands = grid.get_features_in_area(float(uv[m][0]), float(uv[m][1]), radius)
        if not cands:
            out.append(None)
            continue
        best, second, best_idx = U32_MAX, U32_MAX, None
        for k in cands:
            d = descriptor_distance(mp_desc[m], desc[k])
            if d < best:
                second, best, best_idx = best, d, k
            elif d < second:
                second = d
        if best > TH_HIGH:
            out.append(None)
            continue
        if len(cands) > 1 and F32(best) > F32(NN_RATIO * F32(second)):
            out.append(None)
            continue
        out.append(best_idx)
    return out


def search_motion_model(grid, desc, uv, mp_desc, cam, radius=15.0):
    """tracker.rs:1093-1158 -> [kp_idx | None]."""
    out = []
    w, h = 2.0 * cam["cx"], 2.0 * cam["cy"]
    for m in range(len(uv)):
        u, v = float(uv[m][0]), float(uv[m][1])
        if u < 0.0 or u >= w or v < 0.0 or v >= h:
            out.append(None)
            continue
        best, best_idx = U32_MAX, None
        for k in grid.get_features_in_area(u, v, radius):
            d = descriptor_distance(mp_desc[m], desc[k])
            if d < best and d < TH_HIGH:
                best, best_idx = d, k
        out.append(best_idx)
    return out


# ------------------------------------------------------------------------------- local mapping (row f2)
def fuse_search(kp, desc, uvr, mp_desc, desc_threshold=TH_LOW):
    """search_in_neighbors.rs:303-340 over KeyFrame::get_features_in_area (keyframe.rs:408-446, no level bounds)."""
    out = []
    for m in range(len(uvr)):
        u, v, radius = (float(x) for x in uvr[m])
        r2 = radius * radius
        cands = []
        for i in range(len(kp)):
            du, dv = float(kp["x"][i]) - u, float(kp["y"][i]) - v
            if du * du + dv * dv <= r2:
                cands.append(i)
        best, best_idx = U32_MAX, None
        for i in cands:
            d = descriptor_distance(mp_desc[m], desc[i])
            if d < best and d < desc_threshold:
                best, best_idx = d, i
        out.append(best_idx)
    return out


def distinctive_descriptor(descs):
    """map.rs:917-937 (the caller handles 0 / 1 descriptors at :900-915): index of the min-max descriptor."""
    if len(descs) == 0:
        return -1
    if len(descs) == 1:
        return 0
    best_idx, best_max = 0, U32_MAX
    for i in range(len(descs)):
        mx = 0
        for j in range(len(descs)):
            if i != j:
                mx = max(mx, descriptor_distance(descs[i], descs[j]))
        if mx < best_max:
            best_max, best_idx = mx, i
    return best_idx


# --- local_mapping/triangulation.rs ------------------------------------------------------------------------
TRI_CELL = F32(32.0)  # triangulation.rs:339
TRI_MAX_DIM = 64  # :342


def _tri_grid_dims(cam):
    """triangulation.rs:432-436."""
    iw, ih = f64_as_u32(cam["cx"] * 2.0), f64_as_u32(cam["cy"] * 2.0)
    cols = min(f64_as_usize(float(np.ceil(F32(F32(iw) / TRI_CELL)))), TRI_MAX_DIM)
    rows = min(f64_as_usize(float(np.ceil(F32(F32(ih) / TRI_CELL)))), TRI_MAX_DIM)
    return cols, rows


def _tri_build_grid(kp, cols, rows):
    """triangulation.rs:348-363."""
    grid = [[] for _ in range(cols * rows)]
    for idx in range(len(kp)):
        col = min(f64_as_usize(float(F32(F32(kp["x"][idx]) / TRI_CELL))), cols - 1)
        row = min(f64_as_usize(float(F32(F32(kp["y"][idx]) / TRI_CELL))), rows - 1)
        grid[row * cols + col].append(idx)
    return grid


def _tri_candidates(grid, cols, rows, x, y, radius):
    """triangulation.rs:368-392 (f32 arithmetic; `.floor().max(0.0) as usize`, `.ceil() as usize` saturating)."""
    x, y, radius = F32(x), F32(y), F32(radius)
    min_col = f64_as_usize(float(max(np.floor(F32(F32(x - radius) / TRI_CELL)), F32(0.0))))
    max_col = min(f64_as_usize(float(np.ceil(F32(F32(x + radius) / TRI_CELL)))), cols - 1)
    min_row = f64_as_usize(float(max(np.floor(F32(F32(y - radius) / TRI_CELL)), F32(0.0))))
    max_row = min(f64_as_usize(float(np.ceil(F32(F32(y + radius) / TRI_CELL)))), rows - 1)
    out = []
    for row in range(min_row, max_row + 1):
        for col in range(min_col, max_col + 1):
            ci = row * cols + col
            if ci < len(grid):
                out.extend(grid[ci])
    return out


def epipolar_ok(F, x1, y1, x2, y2):
    """triangulation.rs:683-699 given the fundamental matrix F (3x3 f64, computed by the caller exactly as :668-680):
    l2 = F * p1 (nalgebra gemv: column-wise accumulation), dot = (a + b) + c, one rounding per operation."""
    x1, y1, x2, y2 = float(x1), float(y1), float(x2), float(y2)
    l2 = []
    for i in range(3):
        acc = F[i][0] * x1
        acc = acc + F[i][1] * y1
        acc = acc + F[i][2] * 1.0
        l2.append(acc)
    num = abs((l2[0] * x2 + l2[1] * y2) + l2[2] * 1.0)
    den = math.sqrt(l2[0] * l2[0] + l2[1] * l2[1])
    if den < 1e-10:
        return False
    dist = num / den
    return dist * dist < 3.84 * 1.0


def _tri_inner(idx1, cands, kp1, desc1, has_stereo1, kp2, desc2, has_mp2, matched2, F, epipole, max_dist):
    """The body both variants share: triangulation.rs:471-523 == :601-649."""
    best, best_idx2 = max_dist, None
    for idx2 in cands:
        if matched2[idx2] or has_mp2[idx2]:
            continue
        if not has_stereo1[idx1]:
            dx = epipole[0] - float(kp2["x"][idx2])
            dy = epipole[1] - float(kp2["y"][idx2])
            if dx * dx + dy * dy < 100.0 * 1.0:
                continue
        if not epipolar_ok(F, kp1["x"][idx1], kp1["y"][idx1], kp2["x"][idx2], kp2["y"][idx2]):
            continue
        d = descriptor_distance(desc1[idx1], desc2[idx2])
        if d < best and d <= max_dist:
            best_idx2, best = idx2, d
    return best_idx2


def search_for_triangulation(kp1, desc1, has_mp1, has_stereo1, kp2, desc2, has_mp2, F, epipole, cam, max_dist):
    """triangulation.rs:401-530 (spatial-grid variant) -> [(idx1, idx2)] in emission order."""
    cols, rows = _tri_grid_dims(cam)
    grid = _tri_build_grid(kp2, cols, rows)
    matched2 = [False] * len(kp2)
    out = []
    for idx1 in range(len(kp1)):
        if has_mp1[idx1]:
            continue
        cands = _tri_candidates(grid, cols, rows, kp1["x"][idx1], kp1["y"][idx1], 100.0)
        j = _tri_inner(idx1, cands, kp1, desc1, has_stereo1, kp2, desc2, has_mp2, matched2, F, epipole, max_dist)
        if j is not None:
            out.append((idx1, j))
            matched2[j] = True
    return out


def search_for_triangulation_bow(groups, kp1, desc1, has_mp1, has_stereo1, kp2, desc2, has_mp2, F, epipole, max_dist):
    """triangulation.rs:541-658.  groups: [(indices1, indices2)] in the order the caller's HashMap iteration visited the
    vocabulary nodes present in BOTH keyframes (the order is part of the input: the greedy pass depends on it)."""
    matched2 = [False] * len(kp2)
    out = []
    for indices1, indices2 in groups:
        for idx1 in indices1:
            if has_mp1[idx1]:
                continue
            j = _tri_inner(idx1, indices2, kp1, desc1, has_stereo1, kp2, desc2, has_mp2, matched2, F, epipole, max_dist)
            if j is not None:
                out.append((idx1, j))
                matched2[j] = True
    return out
