// oracle/stereo_oracle.cpp — TEST INFRASTRUCTURE ONLY (never linked into / called by the product path).
//
// Line-by-line CPU restatement of the reference's own Rust loops on the stereo front-end path:
//   match_features        /root/reference/src/tracking/frame/stereo.rs:80-161
//   descriptor_distance   /root/reference/src/tracking/frame/stereo.rs:166-175
//   triangulate           /root/reference/src/tracking/frame/stereo.rs:178-216
//   FeatureGrid           /root/reference/src/tracking/tracking_frame.rs:52-128
//   local-map search      /root/reference/src/tracking/tracker.rs:863-923
//   motion-model search   /root/reference/src/tracking/tracker.rs:1093-1158
//   ref-KF BFMatcher      /root/reference/src/tracking/tracker.rs:1001-1010 (cv::BFMatcher(NORM_HAMMING, crossCheck))
//   fusion search         /root/reference/src/local_mapping/search_in_neighbors.rs:303-340 with the linear-scan
//                         KeyFrame::get_features_in_area of src/atlas/map/keyframe.rs:408-446      
//   distinctive descriptor /root/reference/src/atlas/map/map.rs:917-937 (min over i of max_j Hamming)   (row f2)
//   BoW quantisation       /root/reference/src/vocabulary/mod.rs:230-251, 263-275 (tree descent)         (row f2)
//   search_for_triangulation[_bow]  /root/reference/src/local_mapping/triangulation.rs:339-392, 401-530, 541-658,
//                          661-699 (epipolar-gated keyframe <-> keyframe greedy Hamming match)           (row f2)
//
// PARITY UNPINNED for this file: the reference cannot be built here (no Rust toolchain) and holds no test
// for these functions; the only related known-answer vectors are the popcount twins at
// vocabulary/mod.rs:430-441 and loop_closing/corrector.rs:626-633 (checked in tests/test_oracle_golden.py).
// The BFMatcher restatement IS pinned against cv2.BFMatcher in tests.
//
// Rust cast semantics are coded explicitly: float->int `as` saturates (NaN -> 0), int->int `as` wraps.
// Build with -ffp-contract=off so no f64/f32 expression is fused.
#include <algorithm>
#include <cmath>
#include <cstdint>
#include <cstring>
#include <limits>
#include <vector>

extern "C" {

struct OrcKeyPoint {
  float x, y, size, angle, response;
  int octave, class_id;
};
struct OrcDMatch {
  int query_idx, train_idx, img_idx;
  float distance;
};
struct OrcCamera {
  double fx, fy, cx, cy, baseline;
};

// stereo.rs:166-175 — byte-wise popcount of XOR over min(cols) bytes.
uint32_t orc_descriptor_distance(const uint8_t* a, int cols_a, const uint8_t* b, int cols_b) {
  uint32_t d = 0;
  int cols = cols_a < cols_b ? cols_a : cols_b;
  for (int j = 0; j < cols; ++j) d += (uint32_t)__builtin_popcount((unsigned)(a[j] ^ b[j]));
  return d;
}

// stereo.rs:80-161. Returns number of matches (ascending left index).
int orc_match_features(const OrcKeyPoint* kl, const uint8_t* dl, int nl, const OrcKeyPoint* kr, const uint8_t* dr,
                       int nr, const OrcCamera* cam, OrcDMatch* out) {
  const double MIN_DEPTH = 0.1, MAX_DEPTH = 40.0;
  const float VERTICAL_MARGIN = 2.0f;
  const uint32_t TH_HIGH = 100;
  float max_disparity = (float)(cam->fx * cam->baseline / MIN_DEPTH);
  float min_disparity = (float)(cam->fx * cam->baseline / MAX_DEPTH);
  int nm = 0;
  for (int li = 0; li < nl; ++li) {
    float ul = kl[li].x, vl = kl[li].y;
    float min_u = std::fmax(ul - max_disparity, 0.0f);
    float max_u = std::fmin(ul - min_disparity, ((float)nr * ul) / (float)nl);
    uint32_t best = TH_HIGH, second = TH_HIGH;
    int best_idx = -1;
    for (int ri = 0; ri < nr; ++ri) {
      float ur = kr[ri].x, vr = kr[ri].y;
      if (std::fabs(vl - vr) > VERTICAL_MARGIN) continue;
      if (ur < min_u || ur > max_u) continue;
      if (ul <= ur) continue;
      uint32_t d = orc_descriptor_distance(dl + (size_t)li * 32, 32, dr + (size_t)ri * 32, 32);
      if (d < best) {
        second = best;
        best = d;
        best_idx = ri;
      } else if (d < second) {
        second = d;
      }
    }
    if (best_idx >= 0) {
      if ((float)best < 0.9f * (float)second || second == TH_HIGH) {
        out[nm].query_idx = li;
        out[nm].train_idx = best_idx;
        out[nm].img_idx = 0;
        out[nm].distance = (float)best;
        ++nm;
      }
    }
  }
  return nm;
}

// stereo.rs:186-216. points: nl x 3 f64; valid: nl bytes (Option tag).
void orc_triangulate(const OrcKeyPoint* kl, int nl, const OrcKeyPoint* kr, int nr, const OrcDMatch* m, int nm,
                     const OrcCamera* cam, double* points, uint8_t* valid) {
  memset(valid, 0, (size_t)nl);
  for (int i = 0; i < nl * 3; ++i) points[i] = 0.0;
  for (int i = 0; i < nm; ++i) {
    int q = m[i].query_idx, t = m[i].train_idx;
    if (q < 0 || q >= nl || t < 0 || t >= nr) continue;
    double lx = (double)kl[q].x, ly = (double)kl[q].y, rx = (double)kr[t].x;
    double disparity = lx - rx;
    if (std::fabs(disparity) < 0.5) continue;
    double z = cam->fx * cam->baseline / disparity;
    double x = (lx - cam->cx) * z / cam->fx;
    double y = (ly - cam->cy) * z / cam->fy;
    points[3 * q + 0] = x;
    points[3 * q + 1] = y;
    points[3 * q + 2] = z;
    valid[q] = 1;
  }
}

// ---- Rust `as` casts --------------------------------------------------------------------------
static inline uint64_t f64_as_usize(double v) {  // saturating, NaN -> 0
  if (!(v > 0.0)) return 0;
  if (v >= 18446744073709551616.0) return std::numeric_limits<uint64_t>::max();
  return (uint64_t)v;
}
static inline int32_t f64_as_i32(double v) {  // saturating, NaN -> 0
  if (v != v) return 0;
  if (v <= -2147483648.0) return std::numeric_limits<int32_t>::min();
  if (v >= 2147483647.0) return std::numeric_limits<int32_t>::max();
  return (int32_t)v;
}

enum { GRID_COLS = 64, GRID_ROWS = 48 };

// tracking_frame.rs:52-89. cell_start has GRID_COLS*GRID_ROWS+1 entries (CSR); cell_items has n entries:
// row-major cells, ascending keypoint index inside each cell (push order).
void orc_grid_build(const OrcKeyPoint* kp, int n, double img_w, double img_h, int* cell_start, int* cell_items) {
  double inv_w = (double)GRID_COLS / (img_w - 0.0);
  double inv_h = (double)GRID_ROWS / (img_h - 0.0);
  const int nc = GRID_COLS * GRID_ROWS;
  std::vector<int> cell_of(n);
  std::vector<int> count(nc + 1, 0);
  for (int i = 0; i < n; ++i) {
    double x = (double)kp[i].x, y = (double)kp[i].y;
    uint64_t cx = f64_as_usize((x - 0.0) * inv_w);
    uint64_t cy = f64_as_usize((y - 0.0) * inv_h);
    if (cx > GRID_COLS - 1) cx = GRID_COLS - 1;
    if (cy > GRID_ROWS - 1) cy = GRID_ROWS - 1;
    cell_of[i] = (int)(cy * GRID_COLS + cx);
    count[cell_of[i] + 1]++;
  }
  for (int c = 0; c < nc; ++c) count[c + 1] += count[c];
  memcpy(cell_start, count.data(), sizeof(int) * (nc + 1));
  std::vector<int> fill(count.begin(), count.end() - 1);
  for (int i = 0; i < n; ++i) cell_items[fill[cell_of[i]]++] = i;
}

// tracking_frame.rs:104-128 — returns the cell rectangle (inclusive); empty iff x0>x1 or y0>y1.
void orc_grid_query_rect(double x, double y, double r, double img_w, double img_h, int* x0, int* x1, int* y0,
                         int* y1) {
  double inv_w = (double)GRID_COLS / (img_w - 0.0);
  double inv_h = (double)GRID_ROWS / (img_h - 0.0);
  int32_t min_cx = f64_as_i32(std::floor((x - 0.0 - r) * inv_w));
  int32_t max_cx = f64_as_i32(std::ceil((x - 0.0 + r) * inv_w));
  int32_t min_cy = f64_as_i32(std::floor((y - 0.0 - r) * inv_h));
  int32_t max_cy = f64_as_i32(std::ceil((y - 0.0 + r) * inv_h));
  uint64_t ux0 = (uint64_t)(int64_t)(min_cx > 0 ? min_cx : 0);
  uint64_t ux1 = (uint64_t)(int64_t)max_cx;  // i32 -> usize wraps (sign extension)
  if (ux1 > GRID_COLS - 1) ux1 = GRID_COLS - 1;
  uint64_t uy0 = (uint64_t)(int64_t)(min_cy > 0 ? min_cy : 0);
  uint64_t uy1 = (uint64_t)(int64_t)max_cy;
  if (uy1 > GRID_ROWS - 1) uy1 = GRID_ROWS - 1;
  // min_c* can exceed the grid (query right of / below the image): keep as-is, the range is then empty
  *x0 = ux0 > 1000000 ? 1000000 : (int)ux0;
  *x1 = (int)ux1;
  *y0 = uy0 > 1000000 ? 1000000 : (int)uy0;
  *y1 = (int)uy1;
}

// get_features_in_area: writes candidates in scan order; returns count.
int orc_grid_query(const int* cell_start, const int* cell_items, double x, double y, double r, double img_w,
                   double img_h, int* out) {
  int x0, x1, y0, y1, n = 0;
  orc_grid_query_rect(x, y, r, img_w, img_h, &x0, &x1, &y0, &y1);
  for (int cy = y0; cy <= y1; ++cy)
    for (int cx = x0; cx <= x1; ++cx) {
      int c = cy * GRID_COLS + cx;
      for (int k = cell_start[c]; k < cell_start[c + 1]; ++k) out[n++] = cell_items[k];
    }
  return n;
}

// Projection searches. uv: M x 2 f64 (already projected by the host, z > 0 only).
// mode 0 = local-map search (tracker.rs:863-923): accept best <= th_high, ratio test when > 1 candidate.
// mode 1 = motion-model search (tracker.rs:1093-1158): bounds 0<=u<2cx, 0<=v<2cy; d < best && d < th_high.
// out_idx[m] = keypoint index or -1; out_dist[m] = best distance (undefined when -1).
void orc_search_by_projection(const OrcKeyPoint* kp, const uint8_t* desc, int n, double img_w, double img_h,
                              const double* uv, const uint8_t* mp_desc, int M, double radius, uint32_t th_high,
                              float nn_ratio, int mode, double bound_w, double bound_h, int* out_idx,
                              int* out_dist) {
  const int nc = GRID_COLS * GRID_ROWS;
  std::vector<int> cs(nc + 1), items(n > 0 ? n : 1), cand(n > 0 ? n : 1);
  orc_grid_build(kp, n, img_w, img_h, cs.data(), items.data());
  for (int m = 0; m < M; ++m) {
    out_idx[m] = -1;
    out_dist[m] = -1;
    double u = uv[2 * m], v = uv[2 * m + 1];
    if (mode == 1) {
      if (u < 0.0 || u >= bound_w || v < 0.0 || v >= bound_h) continue;
    }
    int nc_ = orc_grid_query(cs.data(), items.data(), u, v, radius, img_w, img_h, cand.data());
    if (mode == 0) {
      if (nc_ == 0) continue;
      uint32_t best = 0xFFFFFFFFu, second = 0xFFFFFFFFu;
      int best_idx = -1;
      for (int k = 0; k < nc_; ++k) {
        uint32_t d = orc_descriptor_distance(mp_desc + (size_t)m * 32, 32, desc + (size_t)cand[k] * 32, 32);
        if (d < best) {
          second = best;
          best = d;
          best_idx = cand[k];
        } else if (d < second) {
          second = d;
        }
      }
      if (best > th_high) continue;
      if (nc_ > 1) {
        if ((float)best > nn_ratio * (float)second) continue;
      }
      out_idx[m] = best_idx;
      out_dist[m] = (int)best;
    } else {
      uint32_t best = 0xFFFFFFFFu;
      int best_idx = -1;
      for (int k = 0; k < nc_; ++k) {
        uint32_t d = orc_descriptor_distance(mp_desc + (size_t)m * 32, 32, desc + (size_t)cand[k] * 32, 32);
        if (d < best && d < th_high) {
          best = d;
          best_idx = cand[k];
        }
      }
      if (best_idx >= 0) {
        out_idx[m] = best_idx;
        out_dist[m] = (int)best;
      }
    }
  }
}

// cv::BFMatcher(NORM_HAMMING, crossCheck=true).match(query, train) as used at tracker.rs:1001-1010
// (query = keyframe descriptors, train = frame descriptors). OpenCV's batchDistance cross-check
// (4.x, pinned against cv2 4.13.0 in tests): sidx[q] = nearest train row of q (first minimum),
// tidx[t] = nearest query row of t (first minimum); the pair (q,t) is kept iff tidx[t]==q and sidx[q]==t.
// Returns matches in ascending query index; distance is the integer Hamming distance as float.
int orc_bf_match_crosscheck(const uint8_t* q, int nq, const uint8_t* t, int nt, OrcDMatch* out) {
  std::vector<int> sidx(nq, -1), sd(nq, 0x7FFFFFFF), tidx(nt, -1);
  for (int i = 0; i < nq; ++i)
    for (int j = 0; j < nt; ++j) {
      int d = (int)orc_descriptor_distance(q + (size_t)i * 32, 32, t + (size_t)j * 32, 32);
      if (d < sd[i]) {
        sd[i] = d;
        sidx[i] = j;
      }
    }
  for (int j = 0; j < nt; ++j) {
    int bd = 0x7FFFFFFF;
    for (int i = 0; i < nq; ++i) {
      int d = (int)orc_descriptor_distance(t + (size_t)j * 32, 32, q + (size_t)i * 32, 32);
      if (d < bd) {
        bd = d;
        tidx[j] = i;
      }
    }
  }
  int n = 0;
  for (int i = 0; i < nq; ++i) {
    int j = sidx[i];
    if (j >= 0 && tidx[j] == i) {
      out[n].query_idx = i;
      out[n].train_idx = j;
      out[n].img_idx = 0;
      out[n].distance = (float)sd[i];
      ++n;
    }
  }
  return n;
}

// search_in_neighbors.rs:303-340: for every (map point, target keyframe) query (u, v, radius) the candidates are ALL
// keyframe keypoints inside the true disc (keyframe.rs:436-443: du*du + dv*dv <= r*r in f64, ascending index); the
// match is the first minimum with dist < desc_threshold.  out_idx[q] = keypoint index or -1.
void orc_fuse_search(const OrcKeyPoint* kp, const uint8_t* desc, int n, const double* uvr, const uint8_t* mp_desc, int M,
                     uint32_t desc_threshold, int* out_idx, int* out_dist) {
  for (int q = 0; q < M; ++q) {
    const double u = uvr[3 * q], v = uvr[3 * q + 1], radius = uvr[3 * q + 2];
    const double radius_sq = radius * radius;
    uint32_t best = 0xFFFFFFFFu;
    int best_idx = -1;
    for (int i = 0; i < n; ++i) {
      double du = (double)kp[i].x - u, dv = (double)kp[i].y - v;
      double dist_sq = du * du + dv * dv;
      if (!(dist_sq <= radius_sq)) continue;
      uint32_t d = orc_descriptor_distance(mp_desc + (size_t)q * 32, 32, desc + (size_t)i * 32, 32);
      if (d < best && d < desc_threshold) {
        best = d;
        best_idx = i;
      }
    }
    out_idx[q] = best_idx;
    out_dist[q] = best_idx >= 0 ? (int)best : -1;
  }
}

// map.rs:917-937: per map point, the descriptor whose maximum distance to the others is smallest (first minimum in
// the given order); one descriptor -> index 0, none -> -1.  offsets: M+1 CSR offsets into desc rows.
void orc_distinctive_descriptors(const uint8_t* desc, const int* offsets, int M, int* out_idx) {
  for (int m = 0; m < M; ++m) {
    const int b = offsets[m], e = offsets[m + 1], k = e - b;
    if (k <= 0) {
      out_idx[m] = -1;
      continue;
    }
    if (k == 1) {
      out_idx[m] = 0;
      continue;
    }
    int best_idx = 0;
    uint32_t best_max = 0xFFFFFFFFu;
    for (int i = 0; i < k; ++i) {
      uint32_t mx = 0;
      for (int j = 0; j < k; ++j)
        if (i != j) {
          uint32_t d = orc_descriptor_distance(desc + (size_t)(b + i) * 32, 32, desc + (size_t)(b + j) * 32, 32);
          if (d > mx) mx = d;
        }
      if (mx < best_max) {
        best_max = mx;
        best_idx = i;
      }
    }
    out_idx[m] = best_idx;
  }
}

// vocabulary/mod.rs:230-251 transform_descriptor + :263-275 get_parent_at_level, per descriptor row as transform()
// (:290-318) calls them.  Tree flattened as in include/vslam_b200.h: node 0 = root, CSR children in VocabNode::children
// order, parent (0xFFFFFFFF = root), word_id (-1 = None -> unwrap_or(0)).
void orc_bow_quantize(const uint8_t* desc, int n, const uint8_t* node_desc, const int* child_begin, const int* child_ids,
                      const uint32_t* parent, const int* word_id, int levels_up, uint32_t* out_word, uint32_t* out_leaf,
                      uint32_t* out_node) {
  for (int i = 0; i < n; ++i) {
    const uint8_t* d = desc + (size_t)i * 32;
    size_t node_id = 0;  // start at root
    int guard = 0;
    while (child_begin[node_id + 1] != child_begin[node_id] && guard++ < 64) {  // !children.is_empty()
      const int b = child_begin[node_id], e = child_begin[node_id + 1];
      int best_child = child_ids[b];
      uint32_t best_dist = orc_descriptor_distance(d, 32, node_desc + (size_t)best_child * 32, 32);
      for (int c = b + 1; c < e; ++c) {
        const int child = child_ids[c];
        const uint32_t dist = orc_descriptor_distance(d, 32, node_desc + (size_t)child * 32, 32);
        if (dist < best_dist) {
          best_dist = dist;
          best_child = child;
        }
      }
      node_id = (size_t)best_child;
    }
    out_word[i] = word_id[node_id] < 0 ? 0u : (uint32_t)word_id[node_id];
    out_leaf[i] = (uint32_t)node_id;
    uint32_t up = (uint32_t)node_id;
    for (int k = 0; k < levels_up; ++k) {
      const uint32_t pa = parent[up];
      if (pa == 0xFFFFFFFFu) break;  // reached root
      up = pa;
    }
    out_node[i] = up;
  }
}

// ---- local_mapping/triangulation.rs: search_for_triangulation (:401-530) and search_for_triangulation_bow (:541-658).
// The pose algebra (epipole :417-428, F = K^-T [t12]x R12 K^-1 :668-680) is loop-invariant nalgebra arithmetic and stays
// with the caller, exactly like the projection of the tracking searches: F (row-major 3x3) and the epipole come in.
// flags1: bit 0 = feature already has a map point, bit 1 = has stereo depth; flags2: bit 0 = has a map point.
struct OrcTriParams {
  double F[9];
  double epipole[2];
  uint32_t max_dist;
  float radius;                        // 100.0f (:440)
  uint32_t image_width, image_height;  // (camera.cx * 2.0) as u32, (camera.cy * 2.0) as u32 (:432-433)
};

static inline uint64_t f32_as_usize(float v) {  // Rust `as usize`: saturating, NaN -> 0
  if (!(v > 0.0f)) return 0;
  if (v >= 18446744073709551616.0f) return ~0ull;
  return (uint64_t)v;
}

// :661-699 check_epipolar_constraint with F given.  nalgebra: l2 = F * p1 accumulates column by column
// ((F_i0 x + F_i1 y) + F_i2 * 1), dot of 3-vectors = (a + b) + c.
static bool orc_epipolar_ok(const double* F, const OrcKeyPoint& k1, const OrcKeyPoint& k2) {
  const double x1 = (double)k1.x, y1 = (double)k1.y, x2 = (double)k2.x, y2 = (double)k2.y;
  double l2[3];
  for (int i = 0; i < 3; ++i) {
    double acc = F[3 * i] * x1;
    acc = acc + F[3 * i + 1] * y1;
    acc = acc + F[3 * i + 2] * 1.0;
    l2[i] = acc;
  }
  const double num = std::fabs((l2[0] * x2 + l2[1] * y2) + l2[2] * 1.0);
  const double den = std::sqrt(l2[0] * l2[0] + l2[1] * l2[1]);
  if (den < 1e-10) return false;
  const double dist = num / den;
  return dist * dist < 3.84 * 1.0;
}

// the candidate loop both variants share (:471-523 == :601-649); returns idx2 or -1
static int orc_tri_best(int idx1, const int* cand, int nc, const OrcKeyPoint* kp1, const uint8_t* desc1,
                        const uint8_t* flags1, const OrcKeyPoint* kp2, const uint8_t* desc2, const uint8_t* flags2,
                        const std::vector<uint8_t>& matched2, const OrcTriParams* P) {
  uint32_t best_dist = P->max_dist;
  int best_idx2 = -1;
  const bool has_stereo1 = (flags1[idx1] & 2) != 0;
  for (int c = 0; c < nc; ++c) {
    const int idx2 = cand[c];
    if (matched2[idx2] || (flags2[idx2] & 1)) continue;
    if (!has_stereo1) {
      const double dx = P->epipole[0] - (double)kp2[idx2].x, dy = P->epipole[1] - (double)kp2[idx2].y;
      if (dx * dx + dy * dy < 100.0 * 1.0) continue;
    }
    if (!orc_epipolar_ok(P->F, kp1[idx1], kp2[idx2])) continue;
    const uint32_t dist = orc_descriptor_distance(desc1 + (size_t)idx1 * 32, 32, desc2 + (size_t)idx2 * 32, 32);
    if (dist < best_dist && dist <= P->max_dist) {
      best_idx2 = idx2;
      best_dist = dist;
    }
  }
  return best_idx2;
}

// Spatial-grid variant.  out_pairs: (idx1, idx2) in emission order; returns the pair count.
int orc_search_for_triangulation(const OrcKeyPoint* kp1, const uint8_t* desc1, const uint8_t* flags1, int n1,
                                 const OrcKeyPoint* kp2, const uint8_t* desc2, const uint8_t* flags2, int n2,
                                 const OrcTriParams* P, int* out_pairs) {
  const float cell = 32.0f;  // GRID_CELL_SIZE (:339)
  const uint64_t max_dim = 64;  // MAX_GRID_DIM (:342)
  const uint64_t cols = std::min(f32_as_usize(std::ceil((float)P->image_width / cell)), max_dim);
  const uint64_t rows = std::min(f32_as_usize(std::ceil((float)P->image_height / cell)), max_dim);
  if (cols == 0 || rows == 0) return 0;  // (the reference would underflow `grid_cols - 1`; not reachable with a real camera)
  std::vector<std::vector<int>> grid(cols * rows);  // build_feature_grid (:348-363)
  for (int i = 0; i < n2; ++i) {
    const uint64_t col = std::min(f32_as_usize(kp2[i].x / cell), cols - 1);
    const uint64_t row = std::min(f32_as_usize(kp2[i].y / cell), rows - 1);
    grid[row * cols + col].push_back(i);
  }
  std::vector<uint8_t> matched2((size_t)std::max(n2, 1), 0);
  std::vector<int> cand;
  int n = 0;
  for (int idx1 = 0; idx1 < n1; ++idx1) {
    if (flags1[idx1] & 1) continue;
    // get_candidates_in_radius (:368-392), f32 arithmetic
    const float x = kp1[idx1].x, y = kp1[idx1].y, r = P->radius;
    const uint64_t min_col = f32_as_usize(std::fmax(std::floor((x - r) / cell), 0.0f));
    const uint64_t max_col = std::min(f32_as_usize(std::ceil((x + r) / cell)), cols - 1);
    const uint64_t min_row = f32_as_usize(std::fmax(std::floor((y - r) / cell), 0.0f));
    const uint64_t max_row = std::min(f32_as_usize(std::ceil((y + r) / cell)), rows - 1);
    cand.clear();
    for (uint64_t row = min_row; row <= max_row; ++row)
      for (uint64_t col = min_col; col <= max_col; ++col) {
        const uint64_t ci = row * cols + col;
        if (ci < grid.size()) cand.insert(cand.end(), grid[ci].begin(), grid[ci].end());
      }
    const int j = orc_tri_best(idx1, cand.data(), (int)cand.size(), kp1, desc1, flags1, kp2, desc2, flags2, matched2, P);
    if (j >= 0) {
      out_pairs[2 * n] = idx1;
      out_pairs[2 * n + 1] = j;
      ++n;
      matched2[j] = 1;
    }
  }
  return n;
}

// FeatureVector variant.  The reference iterates a HashMap<u32, Vec<usize>> (vocabulary/mod.rs:37), so the node order is
// an INPUT: group g = (indices1, indices2) of the g-th visited node present in both keyframes, as CSR lists.
int orc_search_for_triangulation_bow(const OrcKeyPoint* kp1, const uint8_t* desc1, const uint8_t* flags1, int n1,
                                     const OrcKeyPoint* kp2, const uint8_t* desc2, const uint8_t* flags2, int n2,
                                     const OrcTriParams* P, int n_groups, const int* g1_start, const int* g1_items,
                                     const int* g2_start, const int* g2_items, int* out_pairs) {
  std::vector<uint8_t> matched2((size_t)std::max(n2, 1), 0);
  int n = 0;
  for (int g = 0; g < n_groups; ++g)
    for (int a = g1_start[g]; a < g1_start[g + 1]; ++a) {
      const int idx1 = g1_items[a];
      if (idx1 < 0 || idx1 >= n1) continue;  // kp1.get(idx1) -> Err
      if (flags1[idx1] & 1) continue;
      const int j = orc_tri_best(idx1, g2_items + g2_start[g], g2_start[g + 1] - g2_start[g], kp1, desc1, flags1, kp2, desc2,
                                 flags2, matched2, P);
      if (j >= 0) {
        out_pairs[2 * n] = idx1;
        out_pairs[2 * n + 1] = j;
        ++n;
        matched2[j] = 1;
      }
    }
  return n;
}

}  // extern "C"
