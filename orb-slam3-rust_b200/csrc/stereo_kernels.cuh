// Stereo correspondence + depth kernels, bit-exact with the reference's Rust loops:
//   k_match            StereoProcessor::match_features   /root/reference/src/tracking/frame/stereo.rs:80-161
//   k_match_finalize   Vector<DMatch> assembly (ascending left index) + triangulate   stereo.rs:144-157, 186-216
//   k_hamming          descriptor_distance                stereo.rs:166-175
//
// match: one warp per left keypoint scans ALL right keypoints in index order (the reference has no row table);
// right (u,v) and descriptors are staged in shared memory in chunks, descriptors padded to 9 words per row so that
// 32 lanes reading 32 different rows hit 32 different banks.  Hamming = 8 x (xor + popc).  "first index wins" and
// "second best counts duplicates" are reproduced by minimising (dist<<16 | idx) with __reduce_min_sync and taking
// the minimum of everything except the winner as the runner-up.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "vslam_params.h"

namespace vk {

#ifndef VSLAM_FULL
#define VSLAM_FULL 0xffffffffu
#endif

constexpr int kMatchRightChunk = 1024;  // right keypoints staged per pass
constexpr int kMatchLeftPerBlock = 64;  // 8 warps x 8 left keypoints

struct KpXY {  // first 8 bytes of cv::KeyPoint
  float x, y;
};

__device__ __forceinline__ const float* kp_ptr(const FrameLayout& F, int frame, int eye, int i) {
  return reinterpret_cast<const float*>((eye ? F.kp[1] : F.kp[0]) + (long long)frame * F.kp_stride + (size_t)i * 28);
}

// grid = (ceil(cap / 64), frames), block = 256.  ml[frame*cap + li] = (right idx or -1, best distance).
__global__ void __launch_bounds__(256) k_match(const FrameLayout F, const StereoP S, int2* __restrict__ ml) {
  __shared__ float2 s_xy[kMatchRightChunk];
  __shared__ unsigned s_desc[kMatchRightChunk * 9];
  __shared__ unsigned s_key[kMatchLeftPerBlock];
  __shared__ unsigned s_second[kMatchLeftPerBlock];
  const int frame = blockIdx.y, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const ResHdr* hdr = F.hdr + frame;
  const int nL = min(hdr->n_kp[0], F.cap), nR = min(hdr->n_kp[1], F.cap);
  const int l0 = blockIdx.x * kMatchLeftPerBlock;
  if (l0 >= nL) return;
  const unsigned TH = 100u;  // TH_HIGH, stereo.rs:10
  if (tid < kMatchLeftPerBlock) {
    s_key[tid] = (TH << 16) | 0xffffu;
    s_second[tid] = TH;
  }
  const unsigned* descL = reinterpret_cast<const unsigned*>(F.desc[0] + (long long)frame * F.desc_stride);
  const unsigned* descR = reinterpret_cast<const unsigned*>(F.desc[1] + (long long)frame * F.desc_stride);
  for (int r0 = 0; r0 < nR; r0 += kMatchRightChunk) {
    const int rc = min(kMatchRightChunk, nR - r0);
    __syncthreads();
    for (int i = tid; i < rc; i += 256) {
      const float* k = kp_ptr(F, frame, 1, r0 + i);
      s_xy[i] = make_float2(k[0], k[1]);
    }
    for (int i = tid; i < rc * 8; i += 256) s_desc[(i >> 3) * 9 + (i & 7)] = descR[(size_t)r0 * 8 + i];
    __syncthreads();
    for (int q = 0; q < kMatchLeftPerBlock / 8; ++q) {
      const int ls = warp * (kMatchLeftPerBlock / 8) + q;  // slot inside the block
      const int li = l0 + ls;
      if (li >= nL) break;  // warp-uniform
      const float* kl = kp_ptr(F, frame, 0, li);
      const float ul = kl[0], vl = kl[1];
      const float min_u = fmaxf(__fsub_rn(ul, S.max_disparity), 0.0f);
      const float max_u = fminf(__fsub_rn(ul, S.min_disparity), __fdiv_rn(__fmul_rn((float)nR, ul), (float)nL));
      unsigned dl[8];
#pragma unroll
      for (int k = 0; k < 8; ++k) dl[k] = descL[(size_t)li * 8 + k];
      unsigned best = TH, second = TH, bidx = 0xffffu;
      for (int r = lane; r < rc; r += 32) {
        const float2 p = s_xy[r];
        if (fabsf(__fsub_rn(vl, p.y)) > 2.0f) continue;
        if (p.x < min_u || p.x > max_u) continue;
        if (ul <= p.x) continue;
        const unsigned* dr = s_desc + r * 9;
        unsigned d = 0;
#pragma unroll
        for (int k = 0; k < 8; ++k) d += __popc(dl[k] ^ dr[k]);
        if (d < best) {
          second = best;
          best = d;
          bidx = (unsigned)(r0 + r);
        } else if (d < second) {
          second = d;
        }
      }
      const unsigned key = (best << 16) | bidx;
      const unsigned kmin = __reduce_min_sync(VSLAM_FULL, key);
      // runner-up = minimum over everything except the single winning element
      const unsigned other = (key == kmin && best < TH) ? second : best;
      const unsigned smin = __reduce_min_sync(VSLAM_FULL, other);
      if (lane == 0) {
        // merge this chunk's (best, second) summary into the running one (earlier chunk wins ties via idx)
        const unsigned k0 = s_key[ls], s0 = s_second[ls];
        const unsigned b0 = k0 >> 16, b1 = kmin >> 16;
        s_key[ls] = min(k0, kmin);
        s_second[ls] = min(min(s0, smin), max(b0, b1));
      }
    }
  }
  __syncthreads();
  if (tid < kMatchLeftPerBlock) {
    const int li = l0 + tid;
    if (li < nL) {
      const unsigned key = s_key[tid], second = s_second[tid];
      const unsigned best = key >> 16;
      int idx = -1;
      if (best < TH) {  // best_right_idx is Some(..)
        if (__uint2float_rn(best) < __fmul_rn(0.9f, __uint2float_rn(second)) || second == TH) idx = (int)(key & 0xffffu);
      }
      ml[(long long)frame * F.cap + li] = make_int2(idx, (int)best);
    }
  }
}

// grid = frames, block = 256.  Ordered compaction of the per-left results into Vector<DMatch> + triangulation.
__global__ void __launch_bounds__(256) k_match_finalize(const FrameLayout F, const StereoP S,
                                                        const int2* __restrict__ ml) {
  __shared__ int s_warp[8];
  __shared__ int s_base;
  const int frame = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  ResHdr* hdr = F.hdr + frame;
  const int nL = min(hdr->n_kp[0], F.cap);
  int4* matches = reinterpret_cast<int4*>(F.matches + (long long)frame * F.matches_stride);
  double* pts = reinterpret_cast<double*>(F.pts + (long long)frame * F.pts_stride);
  unsigned char* valid = F.valid + (long long)frame * F.valid_stride;
  if (tid == 0) s_base = 0;
  __syncthreads();
  for (int b = 0; b < nL; b += 256) {
    const int li = b + tid;
    int2 m = make_int2(-1, 0);
    if (li < nL) m = ml[(long long)frame * F.cap + li];
    const bool has = m.x >= 0;
    const unsigned bal = __ballot_sync(VSLAM_FULL, has);
    if (lane == 0) s_warp[warp] = __popc(bal);
    __syncthreads();
    int pos = s_base + __popc(bal & ((1u << lane) - 1u));
    for (int w = 0; w < warp; ++w) pos += s_warp[w];
    if (li < nL) {
      double X = 0.0, Y = 0.0, Z = 0.0;
      unsigned char ok = 0;
      if (has) {
        matches[pos] = make_int4(li, m.x, 0, __float_as_int(__int2float_rn(m.y)));
        const float* kl = kp_ptr(F, frame, 0, li);
        const float* kr = kp_ptr(F, frame, 1, m.x);
        const double lx = (double)kl[0], ly = (double)kl[1], rx = (double)kr[0];
        const double disp = __dsub_rn(lx, rx);
        if (!(fabs(disp) < 0.5)) {
          Z = __ddiv_rn(S.fxb, disp);
          X = __ddiv_rn(__dmul_rn(__dsub_rn(lx, S.cx), Z), S.fx);
          Y = __ddiv_rn(__dmul_rn(__dsub_rn(ly, S.cy), Z), S.fy);
          ok = 1;
        }
      }
      pts[3 * (size_t)li + 0] = X;
      pts[3 * (size_t)li + 1] = Y;
      pts[3 * (size_t)li + 2] = Z;
      valid[li] = ok;
    }
    __syncthreads();
    if (tid == 0) {
      int t = 0;
      for (int w = 0; w < 8; ++w) t += s_warp[w];
      s_base += t;
    }
    __syncthreads();
  }
  if (tid == 0) hdr->n_matches = s_base;
}

// ----------------------------------------------------------------------------------------------------------
// Row-bucketed matcher: one 1024-thread block per frame does match_features + DMatch assembly + triangulate.
// The reference scans every right keypoint and keeps those with |vl - vr| <= 2 (stereo.rs:117); the winner is the
// minimum of (distance, right index) and the runner-up the second smallest distance, both independent of the scan
// order.  So the right keypoints are counting-sorted by floor(v) in shared memory and each left keypoint (8 lanes)
// only visits the rows floor(vl - 2.5) .. floor(vl + 2.5) — a superset of the gate, which is then applied exactly.
// Dynamic smem: match_rows_smem_bytes(cap, H).
// ----------------------------------------------------------------------------------------------------------
__host__ __device__ inline size_t match_rows_smem_bytes(int cap, int H) {
  return (size_t)cap * (8 + 36 + 8) + (size_t)((cap + 1) & ~1) * 2 + (size_t)(3 * 1024 + 8) * 4 + 64;
}

__global__ void __launch_bounds__(1024) k_match_rows(const FrameLayout F, const StereoP S, int H) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  __shared__ int s_part[32];
  __shared__ int s_base;
  const int cap = F.cap;
  float2* s_xy = reinterpret_cast<float2*>(smem_raw);                       // [cap] sorted by row
  int2* s_ml = reinterpret_cast<int2*>(s_xy + cap);                         // [cap] per-left result
  unsigned* s_desc = reinterpret_cast<unsigned*>(s_ml + cap);               // [cap][9] sorted by row
  int* s_start = reinterpret_cast<int*>(s_desc + (size_t)cap * 9);          // [3072 + 1] row bucket offsets
  unsigned short* s_idx = reinterpret_cast<unsigned short*>(s_start + 3 * 1024 + 8);  // [cap] original right index
  const int frame = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  ResHdr* hdr = F.hdr + frame;
  const int nL = min(hdr->n_kp[0], cap), nR = min(hdr->n_kp[1], cap);
  const unsigned TH = 100u;  // TH_HIGH, stereo.rs:10
  const int nb = min(H, 3 * 1024);  // row buckets (H <= 3072 enforced by the host)

  // ---- counting sort of the right keypoints by row
  for (int i = tid; i < 3 * 1024 + 1; i += 1024) s_start[i] = 0;
  __syncthreads();
  for (int r = tid; r < nR; r += 1024) {
    const float y = kp_ptr(F, frame, 1, r)[1];
    const int b = min(max((int)floorf(y), 0), nb - 1);
    atomicAdd(&s_start[b], 1);
  }
  __syncthreads();
  {
    const int v0 = s_start[3 * tid], v1 = s_start[3 * tid + 1], v2 = s_start[3 * tid + 2];
    const int sum = v0 + v1 + v2;
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
    const int base = s_part[warp] + incl - sum;
    s_start[3 * tid] = base;
    s_start[3 * tid + 1] = base + v0;
    s_start[3 * tid + 2] = base + v0 + v1;
    if (tid == 1023) s_start[3 * 1024] = base + sum;
  }
  __syncthreads();
  // scatter: s_ml doubles as the running fill pointer table?  No — use a second pass with atomics on a copy kept
  // in the (not yet used) s_ml area: fill[b] = start[b].
  int* fill = reinterpret_cast<int*>(s_ml);  // nb ints <= 2*cap ints?  guarded by the host: 2*cap >= nb
  for (int i = tid; i < nb; i += 1024) fill[i] = s_start[i];
  __syncthreads();
  const unsigned* descR = reinterpret_cast<const unsigned*>(F.desc[1] + (long long)frame * F.desc_stride);
  for (int r = tid; r < nR; r += 1024) {
    const float* k = kp_ptr(F, frame, 1, r);
    const float x = k[0], y = k[1];
    const int b = min(max((int)floorf(y), 0), nb - 1);
    const int p = atomicAdd(&fill[b], 1);
    s_xy[p] = make_float2(x, y);
    s_idx[p] = (unsigned short)r;
    const uint4 d0 = *reinterpret_cast<const uint4*>(descR + (size_t)r * 8);
    const uint4 d1 = *reinterpret_cast<const uint4*>(descR + (size_t)r * 8 + 4);
    unsigned* dp = s_desc + (size_t)p * 9;
    dp[0] = d0.x; dp[1] = d0.y; dp[2] = d0.z; dp[3] = d0.w;
    dp[4] = d1.x; dp[5] = d1.y; dp[6] = d1.z; dp[7] = d1.w;
  }
  __syncthreads();

  // ---- match: 8 lanes per left keypoint
  const unsigned* descL = reinterpret_cast<const unsigned*>(F.desc[0] + (long long)frame * F.desc_stride);
  const int sub = tid & 7;
  // the left keypoint / descriptor of the NEXT round are loaded while the current round is matched (the 10 rounds of a
  // 1200-keypoint frame would otherwise each expose one global-memory round trip: single-pair latency)
  float n_ul = 0.f, n_vl = 0.f;
  uint4 n_a0 = make_uint4(0, 0, 0, 0), n_a1 = make_uint4(0, 0, 0, 0);
  auto fetch_left = [&](int li) {
    if (li < nL) {
      const float* kl = kp_ptr(F, frame, 0, li);
      n_ul = kl[0];
      n_vl = kl[1];
      n_a0 = *reinterpret_cast<const uint4*>(descL + (size_t)li * 8);
      n_a1 = *reinterpret_cast<const uint4*>(descL + (size_t)li * 8 + 4);
    }
  };
  fetch_left(tid >> 3);
  for (int l0 = 0; l0 < nL; l0 += 128) {
    const int li = l0 + (tid >> 3);
    const bool valid = li < nL;
    unsigned key = (TH << 16) | 0xffffu, second = TH;
    const float ul = n_ul, vl = n_vl;
    const uint4 a0 = n_a0, a1 = n_a1;
    fetch_left(li + 128);
    if (valid) {
      const float min_u = fmaxf(__fsub_rn(ul, S.max_disparity), 0.0f);
      const float max_u = fminf(__fsub_rn(ul, S.min_disparity), __fdiv_rn(__fmul_rn((float)nR, ul), (float)nL));
      const int b0 = min(max((int)floorf(vl - 2.5f), 0), nb - 1);
      const int b1 = min(max((int)floorf(vl + 2.5f), 0), nb - 1);
      const int p1 = s_start[b1 + 1];
      for (int p = s_start[b0] + sub; p < p1; p += 8) {
        const float2 q = s_xy[p];
        if (fabsf(__fsub_rn(vl, q.y)) > 2.0f) continue;
        if (q.x < min_u || q.x > max_u) continue;
        if (ul <= q.x) continue;
        const unsigned* dr = s_desc + (size_t)p * 9;
        const unsigned d = __popc(a0.x ^ dr[0]) + __popc(a0.y ^ dr[1]) + __popc(a0.z ^ dr[2]) + __popc(a0.w ^ dr[3]) +
                           __popc(a1.x ^ dr[4]) + __popc(a1.y ^ dr[5]) + __popc(a1.z ^ dr[6]) + __popc(a1.w ^ dr[7]);
        if (d >= TH) continue;  // only d < 100 can change best or second (both start at 100)
        const unsigned k2 = (d << 16) | (unsigned)s_idx[p];
        if (k2 < key) {          // smaller distance, or equal distance with a smaller right index
          second = key >> 16;
          key = k2;
        } else if (d < second) {
          second = d;
        }
      }
    }
    // reduce over the 8 lanes of the group: winner = min key; runner-up = min over everything but the winner
    const unsigned gmask = 0xffu << (lane & 24);  // the 8 lanes that share this left keypoint
    const unsigned kmin = __reduce_min_sync(gmask, key);
    const unsigned other = __reduce_min_sync(gmask, (key == kmin && (key >> 16) < TH) ? second : (key >> 16));
    if (valid && sub == 0) {
      const unsigned best = kmin >> 16;
      int idx = -1;
      if (best < TH) {
        if (__uint2float_rn(best) < __fmul_rn(0.9f, __uint2float_rn(other)) || other == TH) idx = (int)(kmin & 0xffffu);
      }
      s_ml[li] = make_int2(idx, (int)best);
    }
  }
  __syncthreads();

  // ---- Vector<DMatch> in ascending left index + triangulate (stereo.rs:144-157, 186-216)
  int4* matches = reinterpret_cast<int4*>(F.matches + (long long)frame * F.matches_stride);
  double* pts = reinterpret_cast<double*>(F.pts + (long long)frame * F.pts_stride);
  unsigned char* validp = F.valid + (long long)frame * F.valid_stride;
  if (tid == 0) s_base = 0;
  __syncthreads();
  for (int b = 0; b < nL; b += 1024) {
    const int li = b + tid;
    int2 m = make_int2(-1, 0);
    if (li < nL) m = s_ml[li];
    const bool has = m.x >= 0;
    const unsigned bal = __ballot_sync(VSLAM_FULL, has);
    if (lane == 0) s_part[warp] = __popc(bal);
    __syncthreads();
    int pos = s_base + __popc(bal & ((1u << lane) - 1u));
    for (int w = 0; w < warp; ++w) pos += s_part[w];
    if (li < nL) {
      double X = 0.0, Y = 0.0, Z = 0.0;
      unsigned char ok = 0;
      if (has) {
        matches[pos] = make_int4(li, m.x, 0, __float_as_int(__int2float_rn(m.y)));
        const float* kl = kp_ptr(F, frame, 0, li);
        const float* kr = kp_ptr(F, frame, 1, m.x);
        const double lx = (double)kl[0], ly = (double)kl[1], rx = (double)kr[0];
        const double disp = __dsub_rn(lx, rx);
        if (!(fabs(disp) < 0.5)) {
          Z = __ddiv_rn(S.fxb, disp);
          X = __ddiv_rn(__dmul_rn(__dsub_rn(lx, S.cx), Z), S.fx);
          Y = __ddiv_rn(__dmul_rn(__dsub_rn(ly, S.cy), Z), S.fy);
          ok = 1;
        }
      }
      pts[3 * (size_t)li + 0] = X;
      pts[3 * (size_t)li + 1] = Y;
      pts[3 * (size_t)li + 2] = Z;
      validp[li] = ok;
    }
    __syncthreads();
    if (tid == 0) {
      int t = 0;
      for (int w = 0; w < 32; ++w) t += s_part[w];
      s_base += t;
    }
    __syncthreads();
  }
  if (tid == 0) hdr->n_matches = s_base;
}

// descriptor_distance for n pairs: one thread per pair (8 x xor+popc on 32-bit words).
__global__ void k_hamming(const unsigned* __restrict__ a, const unsigned* __restrict__ b, int n,
                          unsigned* __restrict__ out) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  unsigned d = 0;
#pragma unroll
  for (int k = 0; k < 8; ++k) d += __popc(a[(size_t)i * 8 + k] ^ b[(size_t)i * 8 + k]);
  out[i] = d;
}

}  // namespace vk
