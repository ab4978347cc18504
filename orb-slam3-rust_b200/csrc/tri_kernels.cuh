// Keyframe <-> keyframe matching for triangulation, bit-exact with the reference's greedy Rust loops:
//   search_for_triangulation       /root/reference/src/local_mapping/triangulation.rs:401-530 (spatial grid :339-392)
//   search_for_triangulation_bow   triangulation.rs:541-658 (FeatureVector groups)
//   check_epipolar_constraint      triangulation.rs:661-699 (the pose algebra that builds F stays with the caller)
//
// The reference walks the features of keyframe 1 in order and lets each take the best still-unmatched candidate of
// keyframe 2 (`matched2`), so the result depends on the visiting order.  The work splits into
//   k_tri_candidates  (parallel, one warp per visited feature): every candidate that passes the STATIC gates (no map
//                     point, epipole distance, epipolar line distance, Hamming < max_dist) with key = dist << 16 | scan
//                     position, sorted ascending — the greedy loop's choice is always the first still-unmatched entry;
//   k_tri_resolve     (serial in effect: one warp, one lane active at a time) replays the greedy pass over those lists.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "vslam_params.h"

namespace vk {

#ifndef VSLAM_FULL
#define VSLAM_FULL 0xffffffffu
#endif

struct TriP {
  double F[9];
  double ex, ey;       // epipole
  unsigned max_dist;
  float radius;        // 100.0f
  int cols, rows;      // grid dimensions (grid variant), each <= 64
};

constexpr int kTriMaxCells = 64 * 64;

// Rust `as usize` of an f32 (saturating, NaN -> 0), clamped to [0, hi].
__device__ __forceinline__ int f32_as_index(float v, int hi) {
  if (!(v > 0.0f)) return 0;
  if (v >= (float)(hi + 1)) return hi;
  return (int)v;
}

// build_feature_grid (triangulation.rs:348-363) as a CSR: one block of 1024 threads; cell_of: scratch [n].
__global__ void __launch_bounds__(1024) k_tri_grid_build(const unsigned char* __restrict__ kps, int n, int cols, int rows,
                                                         int* __restrict__ cell_start, int* __restrict__ cell_items,
                                                         int* __restrict__ cell_of) {
  __shared__ int s_cnt[kTriMaxCells];
  __shared__ int s_part[32];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int ncell = cols * rows;
  for (int c = tid; c < kTriMaxCells; c += 1024) s_cnt[c] = 0;
  __syncthreads();
  for (int i = tid; i < n; i += 1024) {
    const float* k = reinterpret_cast<const float*>(kps + (size_t)i * 28);
    const int col = f32_as_index(__fdiv_rn(k[0], 32.0f), cols - 1);
    const int row = f32_as_index(__fdiv_rn(k[1], 32.0f), rows - 1);
    const int c = row * cols + col;
    cell_of[i] = c;
    atomicAdd(&s_cnt[c], 1);
  }
  __syncthreads();
  int v[4], sum = 0;
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    v[j] = s_cnt[4 * tid + j];
    sum += v[j];
  }
  int incl = sum;
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
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    s_cnt[4 * tid + j] = base;
    if (4 * tid + j <= ncell) cell_start[4 * tid + j] = base;
    base += v[j];
  }
  if (tid == 1023 && ncell == kTriMaxCells) cell_start[ncell] = base;
  __syncthreads();
  if (warp == 0) {  // stable fill: ascending keypoint index inside a cell == Vec::push order
    for (int b = 0; b < n; b += 32) {
      const int i = b + lane;
      const int c = i < n ? cell_of[i] : -1 - lane;
      const unsigned m = __match_any_sync(VSLAM_FULL, c);
      const int rank = __popc(m & ((1u << lane) - 1u));
      const int start = i < n ? s_cnt[c] : 0;
      __syncwarp();
      if (i < n) {
        cell_items[start + rank] = i;
        if (rank == 0) s_cnt[c] = start + __popc(m);
      }
      __syncwarp();
    }
  }
}

// One warp per visit.  visit_idx1 == nullptr: grid variant, visit v is feature v of keyframe 1 and its candidates are
// the cells get_candidates_in_radius covers; otherwise visit v is feature visit_idx1[v] with the candidates
// g2_items[g2_start[visit_group[v]] .. g2_start[visit_group[v] + 1]).
// Out: cnt[v] = candidates that passed, list[v * cap + k] = (key, idx2) sorted by key for k < min(cnt[v], cap).
__global__ void __launch_bounds__(256) k_tri_candidates(
    const unsigned char* __restrict__ kp1, const unsigned* __restrict__ desc1, const unsigned char* __restrict__ flags1, int n1,
    const unsigned char* __restrict__ kp2, const unsigned* __restrict__ desc2, const unsigned char* __restrict__ flags2, int n2,
    const TriP T, const int* __restrict__ cell_start, const int* __restrict__ cell_items, const int* __restrict__ visit_idx1,
    const int* __restrict__ visit_group, const int* __restrict__ g2_start, const int* __restrict__ g2_items, int n_visits,
    int cap, uint2* __restrict__ scratch, uint2* __restrict__ list, int* __restrict__ cnt) {
  const int v = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  if (v >= n_visits) return;
  const int idx1 = visit_idx1 ? visit_idx1[v] : v;
  int total = 0;
  uint2* raw = scratch + (size_t)v * cap;
  if (idx1 >= 0 && idx1 < n1 && !(flags1[idx1] & 1)) {
    const bool has_stereo1 = (flags1[idx1] & 2) != 0;
    const float* k1 = reinterpret_cast<const float*>(kp1 + (size_t)idx1 * 28);
    const float x1f = k1[0], y1f = k1[1];
    const double x1 = (double)x1f, y1 = (double)y1f;
    // l2 = F * p1 (nalgebra gemv: column by column), den = sqrt(l2.x^2 + l2.y^2)
    double l2[3];
#pragma unroll
    for (int i = 0; i < 3; ++i) {
      double acc = __dmul_rn(T.F[3 * i], x1);
      acc = __dadd_rn(acc, __dmul_rn(T.F[3 * i + 1], y1));
      acc = __dadd_rn(acc, __dmul_rn(T.F[3 * i + 2], 1.0));
      l2[i] = acc;
    }
    const double den = __dsqrt_rn(__dadd_rn(__dmul_rn(l2[0], l2[0]), __dmul_rn(l2[1], l2[1])));
    unsigned d1[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) d1[k] = desc1[(size_t)idx1 * 8 + k];
    const unsigned lt = (1u << lane) - 1u;
    int pos0 = 0;  // scan position of the first candidate of the current span
    auto span = [&](const int* items, int s, int e) {
      for (int k0 = s; k0 < e; k0 += 32) {
        const int k = k0 + lane;
        bool ok = false;
        unsigned key = 0;
        int idx2 = 0;
        if (k < e) {
          idx2 = items[k];
          ok = idx2 >= 0 && idx2 < n2 && !(flags2[idx2] & 1);
          if (ok) {
            const float* k2 = reinterpret_cast<const float*>(kp2 + (size_t)idx2 * 28);
            const double x2 = (double)k2[0], y2 = (double)k2[1];
            if (!has_stereo1) {
              const double dx = __dsub_rn(T.ex, x2), dy = __dsub_rn(T.ey, y2);
              if (__dadd_rn(__dmul_rn(dx, dx), __dmul_rn(dy, dy)) < 100.0) ok = false;
            }
            if (ok) {
              const double num =
                  fabs(__dadd_rn(__dadd_rn(__dmul_rn(l2[0], x2), __dmul_rn(l2[1], y2)), __dmul_rn(l2[2], 1.0)));
              if (den < 1e-10) ok = false;
              else {
                const double dist = __ddiv_rn(num, den);
                ok = __dmul_rn(dist, dist) < 3.84;
              }
            }
            if (ok) {
              const uint4* dp = reinterpret_cast<const uint4*>(desc2 + (size_t)idx2 * 8);
              const uint4 a = dp[0], b = dp[1];
              const unsigned d = __popc(d1[0] ^ a.x) + __popc(d1[1] ^ a.y) + __popc(d1[2] ^ a.z) + __popc(d1[3] ^ a.w) +
                                 __popc(d1[4] ^ b.x) + __popc(d1[5] ^ b.y) + __popc(d1[6] ^ b.z) + __popc(d1[7] ^ b.w);
              ok = d < T.max_dist;  // best_dist starts at max_dist: only dist < max_dist can ever win
              key = (d << 16) | (unsigned)(pos0 + (k - s));
            }
          }
        }
        const unsigned m = __ballot_sync(VSLAM_FULL, ok);
        const int p = total + __popc(m & lt);
        if (ok && p < cap) raw[p] = make_uint2(key, (unsigned)idx2);
        total += __popc(m);
      }
      pos0 += e - s;
    };
    if (!visit_idx1) {
      // get_candidates_in_radius (triangulation.rs:368-392), f32
      const float r = T.radius;
      const int min_col = f32_as_index(fmaxf(floorf(__fdiv_rn(__fsub_rn(x1f, r), 32.0f)), 0.0f), 1 << 30);
      const int max_col = f32_as_index(ceilf(__fdiv_rn(__fadd_rn(x1f, r), 32.0f)), T.cols - 1);
      const int min_row = f32_as_index(fmaxf(floorf(__fdiv_rn(__fsub_rn(y1f, r), 32.0f)), 0.0f), 1 << 30);
      const int max_row = f32_as_index(ceilf(__fdiv_rn(__fadd_rn(y1f, r), 32.0f)), T.rows - 1);
      if (min_col <= max_col)
        for (int row = min_row; row <= max_row; ++row)
          span(cell_items, cell_start[row * T.cols + min_col], cell_start[row * T.cols + max_col + 1]);
    } else {
      const int g = visit_group[v];
      span(g2_items, g2_start[g], g2_start[g + 1]);
    }
  }
  __syncwarp();
  // rank sort (keys are unique: the scan position is part of the key); lists are short
  const int c = min(total, cap);
  uint2* out = list + (size_t)v * cap;
  for (int i = lane; i < c; i += 32) {
    const uint2 me = raw[i];
    int rank = 0;
    for (int j = 0; j < c; ++j) rank += raw[j].x < me.x;
    out[rank] = me;
  }
  if (lane == 0) cnt[v] = total;
}

// Greedy replay: visits in order, each takes its first still-unmatched candidate.  One warp; lanes prefetch the head of
// 32 lists at a time, then resolve one after the other.  out_pairs (idx1, idx2) in emission order; n_pairs[0] = count,
// n_pairs[1] = largest candidate count (> cap means the lists were truncated: the host re-runs with a larger cap).
__global__ void __launch_bounds__(32) k_tri_resolve(const uint2* __restrict__ list, const int* __restrict__ cnt, int cap,
                                                    const int* __restrict__ visit_idx1, int n_visits, int n2,
                                                    int* __restrict__ out_pairs, int* __restrict__ n_pairs) {
  extern __shared__ unsigned s_matched[];  // bitmap over keyframe-2 features
  const int lane = threadIdx.x;
  for (int i = lane; i < (n2 + 31) / 32; i += 32) s_matched[i] = 0;
  __syncwarp();
  int n_out = 0, max_cnt = 0;
  for (int base = 0; base < n_visits; base += 32) {
    const int v = base + lane;
    const int c = v < n_visits ? cnt[v] : 0;
    max_cnt = max(max_cnt, c);
    const int cc = min(c, cap);
    const uint2 head = cc > 0 ? list[(size_t)v * cap] : make_uint2(0, 0);
    const unsigned any = __ballot_sync(VSLAM_FULL, cc > 0);
    for (unsigned rest = any; rest; rest &= rest - 1) {
      const int l = __ffs(rest) - 1;
      int res = -1;
      if (lane == l) {
        for (int k = 0; k < cc; ++k) {
          const unsigned idx2 = k == 0 ? head.y : list[(size_t)v * cap + k].y;
          if (!((s_matched[idx2 >> 5] >> (idx2 & 31)) & 1u)) {
            s_matched[idx2 >> 5] |= 1u << (idx2 & 31);
            res = (int)idx2;
            break;
          }
        }
        if (res >= 0) {
          out_pairs[2 * n_out] = visit_idx1 ? visit_idx1[v] : v;
          out_pairs[2 * n_out + 1] = res;
        }
      }
      __syncwarp();
      n_out += __shfl_sync(VSLAM_FULL, res, l) >= 0 ? 1 : 0;
    }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) max_cnt = max(max_cnt, __shfl_xor_sync(VSLAM_FULL, max_cnt, o));
  if (lane == 0) {
    n_pairs[0] = n_out;
    n_pairs[1] = max_cnt;
  }
}

}  // namespace vk
