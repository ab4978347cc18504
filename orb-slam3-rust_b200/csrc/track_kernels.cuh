// Tracking-side kernels, bit-exact (bug-compatible) with the reference's Rust loops:
//   k_grid_build        FeatureGrid::new                    /root/reference/src/tracking/tracking_frame.rs:52-89
//   k_grid_count/fill   FeatureGrid::get_features_in_area   tracking_frame.rs:104-128
//   k_search            the two projection searches in Tracker   src/tracking/tracker.rs:863-923, 1093-1158
//   k_nearest/k_cross   BFMatcher(NORM_HAMMING, crossCheck=true) src/tracking/tracker.rs:1001-1010
//   k_fuse_search       fusion search of local mapping          src/local_mapping/search_in_neighbors.rs:303-340
//   k_distinctive       Map::compute_distinctive_descriptors    src/atlas/map/map.rs:917-937
//   k_bow_quantize      OrbVocabulary::transform_descriptor     src/vocabulary/mod.rs:230-251, 263-275
//
// Rust cast semantics reproduced explicitly: f64 -> usize / i32 `as` saturates (NaN -> 0); i32 -> usize `as`
// sign-extends, so a negative max cell clamps to the LAST column/row (SURVEY.md §8.C trap 6).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "vslam_params.h"

namespace vk {

#ifndef VSLAM_FULL
#define VSLAM_FULL 0xffffffffu
#endif

constexpr int kGridCols = 64, kGridRows = 48, kGridCells = kGridCols * kGridRows;

struct GridP {
  double inv_w, inv_h;  // GRID_COLS / (img_w - 0.0), GRID_ROWS / (img_h - 0.0)
};

struct SearchP {
  int mode;  // 0 local-map, 1 motion-model
  double radius;
  unsigned th_high;
  float nn_ratio;
  double bound_w, bound_h;
};

__device__ __forceinline__ int cell_of_coord(double x, double inv, int ncell) {
  // ((x - min) * inv) as usize, then .min(ncell - 1)
  double v = __dmul_rn(__dsub_rn(x, 0.0), inv);
  if (!(v > 0.0)) return 0;  // negative or NaN saturate to 0
  if (v >= (double)ncell) return ncell - 1;
  return (int)v;
}

__device__ __forceinline__ int f64_as_i32(double v) {
  if (v != v) return 0;
  if (v <= -2147483648.0) return INT_MIN;
  if (v >= 2147483647.0) return INT_MAX;
  return (int)v;
}

// inclusive cell rectangle of get_features_in_area(x, y, r); empty iff x0 > x1 or y0 > y1.
__device__ __forceinline__ void query_rect(const GridP& G, double x, double y, double r, int& x0, int& x1, int& y0,
                                           int& y1) {
  int min_cx = f64_as_i32(floor(__dmul_rn(__dsub_rn(__dsub_rn(x, 0.0), r), G.inv_w)));
  int max_cx = f64_as_i32(ceil(__dmul_rn(__dadd_rn(__dsub_rn(x, 0.0), r), G.inv_w)));
  int min_cy = f64_as_i32(floor(__dmul_rn(__dsub_rn(__dsub_rn(y, 0.0), r), G.inv_h)));
  int max_cy = f64_as_i32(ceil(__dmul_rn(__dadd_rn(__dsub_rn(y, 0.0), r), G.inv_h)));
  x0 = max(min_cx, 0);
  x1 = max_cx < 0 ? kGridCols - 1 : min(max_cx, kGridCols - 1);  // (i32 as usize).min(cols-1): negative wraps high
  y0 = max(min_cy, 0);
  y1 = max_cy < 0 ? kGridRows - 1 : min(max_cy, kGridRows - 1);
}

// One block (1024 threads) builds the CSR grid of one keypoint set.
// kps: cv::KeyPoint records (28 B); cell_start[kGridCells + 1]; cell_items[n]; cell_of: scratch [n].
__global__ void __launch_bounds__(1024) k_grid_build(const unsigned char* __restrict__ kps, int n, GridP G,
                                                     int* __restrict__ cell_start, int* __restrict__ cell_items,
                                                     int* __restrict__ cell_of) {
  __shared__ int s_cnt[kGridCells];
  __shared__ int s_part[32];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  for (int c = tid; c < kGridCells; c += 1024) s_cnt[c] = 0;
  __syncthreads();
  for (int i = tid; i < n; i += 1024) {
    const float* k = reinterpret_cast<const float*>(kps + (size_t)i * 28);
    int cx = cell_of_coord((double)k[0], G.inv_w, kGridCols);
    int cy = cell_of_coord((double)k[1], G.inv_h, kGridRows);
    int c = cy * kGridCols + cx;
    cell_of[i] = c;
    atomicAdd(&s_cnt[c], 1);
  }
  __syncthreads();
  // exclusive scan of 3072 counts: 3 per thread
  int v0 = s_cnt[3 * tid], v1 = s_cnt[3 * tid + 1], v2 = s_cnt[3 * tid + 2];
  int sum = v0 + v1 + v2, incl = sum;
#pragma unroll
  for (int d = 1; d < 32; d <<= 1) {
    int o = __shfl_up_sync(VSLAM_FULL, incl, d);
    if (lane >= d) incl += o;
  }
  if (lane == 31) s_part[warp] = incl;
  __syncthreads();
  if (warp == 0) {
    int p = s_part[lane], pi = p;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
      int o = __shfl_up_sync(VSLAM_FULL, pi, d);
      if (lane >= d) pi += o;
    }
    s_part[lane] = pi - p;
  }
  __syncthreads();
  int base = s_part[warp] + incl - sum;
  __syncthreads();
  s_cnt[3 * tid] = base;
  s_cnt[3 * tid + 1] = base + v0;
  s_cnt[3 * tid + 2] = base + v0 + v1;
  cell_start[3 * tid] = base;
  cell_start[3 * tid + 1] = base + v0;
  cell_start[3 * tid + 2] = base + v0 + v1;
  if (tid == 1023) cell_start[kGridCells] = base + sum;
  __syncthreads();
  // stable fill (ascending keypoint index inside a cell == Vec::push order): one warp walks the keypoints in order
  if (warp == 0) {
    for (int b = 0; b < n; b += 32) {
      int i = b + lane;
      int c = i < n ? cell_of[i] : -1 - lane;  // distinct dummies so padding lanes never group
      unsigned m = __match_any_sync(VSLAM_FULL, c);
      int rank = __popc(m & ((1u << lane) - 1u));
      int start = i < n ? s_cnt[c] : 0;
      __syncwarp();
      if (i < n) {
        cell_items[start + rank] = i;
        if (rank == 0) s_cnt[c] = start + __popc(m);
      }
      __syncwarp();
    }
  }
}

// Number of candidates of each query (thread per query).  xyr: nq x (x, y, r).
__global__ void k_grid_count(const int* __restrict__ cell_start, GridP G, const double* __restrict__ xyr, int nq,
                             int* __restrict__ counts) {
  int q = blockIdx.x * blockDim.x + threadIdx.x;
  if (q >= nq) return;
  int x0, x1, y0, y1, c = 0;
  query_rect(G, xyr[3 * q], xyr[3 * q + 1], xyr[3 * q + 2], x0, x1, y0, y1);
  if (x0 <= x1)
    for (int cy = y0; cy <= y1; ++cy) c += cell_start[cy * kGridCols + x1 + 1] - cell_start[cy * kGridCols + x0];
  counts[q] = c;
}

// Single-block exclusive scan of nq counts into out_start[nq + 1].
__global__ void __launch_bounds__(1024) k_scan_counts(const int* __restrict__ counts, int nq, int* __restrict__ out_start) {
  __shared__ int s_part[32];
  __shared__ int s_run;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  if (tid == 0) s_run = 0;
  __syncthreads();
  for (int b = 0; b < nq; b += 1024) {
    int i = b + tid;
    int v = i < nq ? counts[i] : 0, incl = v;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
      int o = __shfl_up_sync(VSLAM_FULL, incl, d);
      if (lane >= d) incl += o;
    }
    if (lane == 31) s_part[warp] = incl;
    __syncthreads();
    if (warp == 0) {
      int p = s_part[lane], pi = p;
#pragma unroll
      for (int d = 1; d < 32; d <<= 1) {
        int o = __shfl_up_sync(VSLAM_FULL, pi, d);
        if (lane >= d) pi += o;
      }
      s_part[lane] = pi - p;
    }
    __syncthreads();
    int excl = s_run + s_part[warp] + incl - v;
    if (i < nq) out_start[i] = excl;
    __syncthreads();
    if (tid == 1023) s_run = excl + v;
    __syncthreads();
  }
  if (tid == 0) out_start[nq] = s_run;
}

// Candidate lists in scan order (row-major cells, ascending index inside a cell); warp per query.
__global__ void k_grid_fill(const int* __restrict__ cell_start, const int* __restrict__ cell_items, GridP G,
                            const double* __restrict__ xyr, int nq, const int* __restrict__ out_start,
                            int* __restrict__ out_items, int out_capacity) {
  int q = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  if (q >= nq) return;
  int x0, x1, y0, y1;
  query_rect(G, xyr[3 * q], xyr[3 * q + 1], xyr[3 * q + 2], x0, x1, y0, y1);
  if (x0 > x1) return;
  int o = out_start[q];
  for (int cy = y0; cy <= y1; ++cy) {
    int s = cell_start[cy * kGridCols + x0], e = cell_start[cy * kGridCols + x1 + 1];
    for (int k = s + lane; k < e; k += 32)
      if (o + (k - s) < out_capacity) out_items[o + (k - s)] = cell_items[k];
    o += e - s;
  }
}

// Projection search: one warp per map point.  desc: frame descriptors [n][32]; uv: M x 2; mp_desc [M][32].
__global__ void __launch_bounds__(256) k_search(const int* __restrict__ cell_start, const int* __restrict__ cell_items,
                                                const unsigned* __restrict__ desc, GridP G, SearchP S,
                                                const double* __restrict__ uv, const unsigned* __restrict__ mp_desc,
                                                int M, int* __restrict__ out_idx, int* __restrict__ out_dist) {
  const int m = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  if (m >= M) return;
  const double u = uv[2 * m], v = uv[2 * m + 1];
  int res_idx = -1, res_dist = -1;
  bool in_bounds = true;
  if (S.mode == 1) in_bounds = !(u < 0.0 || u >= S.bound_w || v < 0.0 || v >= S.bound_h);
  if (in_bounds) {
    int x0, x1, y0, y1;
    query_rect(G, u, v, S.radius, x0, x1, y0, y1);
    unsigned md[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) md[k] = mp_desc[(size_t)m * 8 + k];
    const unsigned INF = 0x7ffu;
    unsigned best = INF, second = INF, bpos = 0xfffffu;
    int bidx = -1, ncand = 0;
    if (x0 <= x1) {
      for (int cy = y0; cy <= y1; ++cy) {
        const int s = cell_start[cy * kGridCols + x0], e = cell_start[cy * kGridCols + x1 + 1];
        for (int k = s + lane; k < e; k += 32) {
          const int kp = cell_items[k];
          const uint4* dp = reinterpret_cast<const uint4*>(desc + (size_t)kp * 8);
          const uint4 a = dp[0], b = dp[1];
          unsigned d = __popc(md[0] ^ a.x) + __popc(md[1] ^ a.y) + __popc(md[2] ^ a.z) + __popc(md[3] ^ a.w) +
                       __popc(md[4] ^ b.x) + __popc(md[5] ^ b.y) + __popc(md[6] ^ b.z) + __popc(md[7] ^ b.w);
          if (S.mode == 1 && !(d < S.th_high)) continue;  // motion model: only d < TH_HIGH competes
          if (d < best) {
            second = best;
            best = d;
            bpos = (unsigned)(ncand + (k - s));
            bidx = kp;
          } else if (d < second) {
            second = d;
          }
        }
        ncand += e - s;
      }
    }
    const unsigned key = (best << 20) | bpos;
    const unsigned kmin = __reduce_min_sync(VSLAM_FULL, key);
    const unsigned gbest = kmin >> 20;
    const unsigned other = (key == kmin && best != INF) ? second : best;
    const unsigned gsecond = __reduce_min_sync(VSLAM_FULL, other);
    const unsigned win = __ballot_sync(VSLAM_FULL, key == kmin && best != INF);
    if (win) {
      const int widx = __shfl_sync(VSLAM_FULL, bidx, __ffs(win) - 1);
      if (S.mode == 0) {
        bool ok = gbest <= S.th_high;
        if (ok && ncand > 1) ok = !(__uint2float_rn(gbest) > __fmul_rn(S.nn_ratio, __uint2float_rn(gsecond)));
        if (ok) {
          res_idx = widx;
          res_dist = (int)gbest;
        }
      } else {
        res_idx = widx;
        res_dist = (int)gbest;
      }
    }
  }
  if (lane == 0) {
    out_idx[m] = res_idx;
    out_dist[m] = res_dist;
  }
}

// Nearest row of B for every row of A (first minimum): warp per A row, B rows staged in shared memory.
constexpr int kNearChunk = 1024;
__global__ void __launch_bounds__(256) k_nearest(const unsigned* __restrict__ A, int nA, const unsigned* __restrict__ B,
                                                 int nB, int* __restrict__ out_idx, int* __restrict__ out_dist) {
  __shared__ unsigned s_desc[kNearChunk * 9];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int a = blockIdx.x * 8 + warp;
  unsigned da[8];
#pragma unroll
  for (int k = 0; k < 8; ++k) da[k] = a < nA ? A[(size_t)a * 8 + k] : 0u;
  unsigned long long best = ~0ull;
  for (int r0 = 0; r0 < nB; r0 += kNearChunk) {
    const int rc = min(kNearChunk, nB - r0);
    __syncthreads();
    for (int i = tid; i < rc * 8; i += 256) s_desc[(i >> 3) * 9 + (i & 7)] = B[(size_t)r0 * 8 + i];
    __syncthreads();
    if (a < nA) {
      for (int r = lane; r < rc; r += 32) {
        const unsigned* dr = s_desc + r * 9;
        unsigned d = 0;
#pragma unroll
        for (int k = 0; k < 8; ++k) d += __popc(da[k] ^ dr[k]);
        unsigned long long key = ((unsigned long long)d << 32) | (unsigned)(r0 + r);
        best = key < best ? key : best;
      }
    }
  }
  if (a < nA) {
    unsigned hi = (unsigned)(best >> 32), lo = (unsigned)best;
    // 64-bit min via two 32-bit reductions: first the distance, then the index among the minimisers
    unsigned dmin = __reduce_min_sync(VSLAM_FULL, hi);
    unsigned imin = __reduce_min_sync(VSLAM_FULL, hi == dmin ? lo : 0xffffffffu);
    if (lane == 0) {
      out_idx[a] = nB > 0 ? (int)imin : -1;
      out_dist[a] = nB > 0 ? (int)dmin : 0x7fffffff;
    }
  }
}

// Cross-check + ordered compaction (single block): (q, sidx[q]) kept iff tidx[sidx[q]] == q.
__global__ void __launch_bounds__(256) k_cross(const int* __restrict__ sidx, const int* __restrict__ sdist,
                                               const int* __restrict__ tidx, int nq, int4* __restrict__ matches,
                                               int* __restrict__ n_out) {
  __shared__ int s_warp[8];
  __shared__ int s_base;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  if (tid == 0) s_base = 0;
  __syncthreads();
  for (int b = 0; b < nq; b += 256) {
    const int q = b + tid;
    int j = q < nq ? sidx[q] : -1;
    const bool has = j >= 0 && tidx[j] == q;
    const unsigned bal = __ballot_sync(VSLAM_FULL, has);
    if (lane == 0) s_warp[warp] = __popc(bal);
    __syncthreads();
    int pos = s_base + __popc(bal & ((1u << lane) - 1u));
    for (int w = 0; w < warp; ++w) pos += s_warp[w];
    if (has) matches[pos] = make_int4(q, j, 0, __float_as_int(__int2float_rn(sdist[q])));
    __syncthreads();
    if (tid == 0) {
      int t = 0;
      for (int w = 0; w < 8; ++w) t += s_warp[w];
      s_base += t;
    }
    __syncthreads();
  }
  if (tid == 0) *n_out = s_base;
}

// Fusion search: one warp per query, lanes stride over ALL keyframe keypoints (the reference's linear scan).
// kps: cv::KeyPoint records (28 B); desc [n][32]; uvr [M][3] f64 (u, v, radius).
__global__ void __launch_bounds__(256) k_fuse_search(const unsigned char* __restrict__ kps,
                                                     const unsigned* __restrict__ desc, int n,
                                                     const double* __restrict__ uvr, const unsigned* __restrict__ mp_desc,
                                                     int M, unsigned desc_threshold, int* __restrict__ out_idx,
                                                     int* __restrict__ out_dist) {
  const int q = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  if (q >= M) return;
  const double u = uvr[3 * q], v = uvr[3 * q + 1], radius = uvr[3 * q + 2];
  const double radius_sq = __dmul_rn(radius, radius);
  unsigned md[8];
#pragma unroll
  for (int k = 0; k < 8; ++k) md[k] = mp_desc[(size_t)q * 8 + k];
  unsigned long long best = ~0ull;  // (dist << 32) | index: first minimum == smallest key
  for (int i = lane; i < n; i += 32) {
    const float* k = reinterpret_cast<const float*>(kps + (size_t)i * 28);
    const double du = __dsub_rn((double)k[0], u), dv = __dsub_rn((double)k[1], v);
    const double dist_sq = __dadd_rn(__dmul_rn(du, du), __dmul_rn(dv, dv));
    if (!(dist_sq <= radius_sq)) continue;
    const uint4* dp = reinterpret_cast<const uint4*>(desc + (size_t)i * 8);
    const uint4 a = dp[0], b = dp[1];
    const unsigned d = __popc(md[0] ^ a.x) + __popc(md[1] ^ a.y) + __popc(md[2] ^ a.z) + __popc(md[3] ^ a.w) +
                       __popc(md[4] ^ b.x) + __popc(md[5] ^ b.y) + __popc(md[6] ^ b.z) + __popc(md[7] ^ b.w);
    if (d >= desc_threshold) continue;
    const unsigned long long key = ((unsigned long long)d << 32) | (unsigned)i;
    best = key < best ? key : best;
  }
  const unsigned hi = (unsigned)(best >> 32), lo = (unsigned)best;
  const unsigned dmin = __reduce_min_sync(VSLAM_FULL, hi);
  const unsigned imin = __reduce_min_sync(VSLAM_FULL, hi == dmin ? lo : 0xffffffffu);
  if (lane == 0) {
    const bool found = dmin != 0xffffffffu;
    out_idx[q] = found ? (int)imin : -1;
    out_dist[q] = found ? (int)dmin : -1;
  }
}

// Distinctive descriptor: one warp per map point; for each i the lanes take max_j d(i, j), lane 0 keeps the first
// minimum of those maxima.
__global__ void __launch_bounds__(256) k_distinctive(const unsigned* __restrict__ desc, const int* __restrict__ offsets,
                                                     int M, int* __restrict__ out_idx) {
  const int m = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  if (m >= M) return;
  const int b = offsets[m], k = offsets[m + 1] - b;
  int best_idx = k <= 0 ? -1 : 0;
  if (k > 1) {
    unsigned best_max = 0xffffffffu;
    for (int i = 0; i < k; ++i) {
      const uint4* pi = reinterpret_cast<const uint4*>(desc + (size_t)(b + i) * 8);
      const uint4 a0 = pi[0], a1 = pi[1];
      unsigned mx = 0;
      for (int j = lane; j < k; j += 32) {
        if (j == i) continue;
        const uint4* pj = reinterpret_cast<const uint4*>(desc + (size_t)(b + j) * 8);
        const uint4 c0 = pj[0], c1 = pj[1];
        const unsigned d = __popc(a0.x ^ c0.x) + __popc(a0.y ^ c0.y) + __popc(a0.z ^ c0.z) + __popc(a0.w ^ c0.w) +
                           __popc(a1.x ^ c1.x) + __popc(a1.y ^ c1.y) + __popc(a1.z ^ c1.z) + __popc(a1.w ^ c1.w);
        mx = max(mx, d);
      }
      mx = __reduce_max_sync(VSLAM_FULL, mx);
      if (mx < best_max) {
        best_max = mx;
        best_idx = i;
      }
    }
  }
  if (lane == 0) out_idx[m] = best_idx;
}

// ----------------------------------------------------------------------------------------------------------
// BoW quantisation: OrbVocabulary::transform_descriptor + get_parent_at_level,
// /root/reference/src/vocabulary/mod.rs:230-251, 263-275 (called per descriptor row by transform(), :290-318).
// 8 lanes per descriptor (one 32-bit word each, coalesced 32-byte node descriptors), 4 descriptors per warp.  At each
// node the children are scanned in list order; the closest child wins, first one on ties (strict `<`).  A node
// without children is a leaf; word = word_id or 0 (`unwrap_or(0)`).  The walk is capped at kBowMaxDepth levels (the
// reference would not terminate on a cyclic child list).
// ----------------------------------------------------------------------------------------------------------
constexpr int kBowMaxDepth = 64;
__global__ void __launch_bounds__(256) k_bow_quantize(const unsigned* __restrict__ desc, int n,
                                                      const unsigned* __restrict__ node_desc,
                                                      const int* __restrict__ child_begin,
                                                      const int* __restrict__ child_ids,
                                                      const unsigned* __restrict__ parent,
                                                      const int* __restrict__ word_id, int levels_up,
                                                      unsigned* __restrict__ out_word, unsigned* __restrict__ out_leaf,
                                                      unsigned* __restrict__ out_node) {
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  const int g = t >> 3, sub = t & 7, lane = threadIdx.x & 31;
  if (g >= n) return;  // whole 8-lane groups leave together
  const unsigned gmask = 0xffu << (lane & 24);
  const unsigned mine = desc[(size_t)g * 8 + sub];
  int node = 0;
  for (int depth = 0; depth < kBowMaxDepth; ++depth) {
    const int b = child_begin[node], e = child_begin[node + 1];
    if (b == e) break;
    unsigned best_dist = 0xffffffffu;
    int best_child = node;
    for (int c = b; c < e; ++c) {
      const int child = child_ids[c];
      unsigned d = (unsigned)__popc(mine ^ node_desc[(size_t)child * 8 + sub]);
      d += __shfl_xor_sync(gmask, d, 1);
      d += __shfl_xor_sync(gmask, d, 2);
      d += __shfl_xor_sync(gmask, d, 4);
      if (c == b || d < best_dist) {
        best_dist = d;
        best_child = child;
      }
    }
    node = best_child;
  }
  if (sub == 0) {
    const int w = word_id[node];
    out_word[g] = w < 0 ? 0u : (unsigned)w;
    out_leaf[g] = (unsigned)node;
    unsigned up = (unsigned)node;
    for (int i = 0; i < levels_up; ++i) {
      const unsigned pa = parent[up];
      if (pa == 0xffffffffu) break;
      up = pa;
    }
    out_node[g] = up;
  }
}

}  // namespace vk
