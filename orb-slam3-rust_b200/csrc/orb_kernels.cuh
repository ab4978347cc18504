// Hand-written sm_100a kernels for ORB extraction, bit-exact with cv::ORB as the reference configures it
// (/root/reference/src/tracking/frame/stereo.rs:38-48,72-73; arithmetic spec: SURVEY.md §8.A0-A6).
//
//   k_resize   A1  INTER_LINEAR_EXACT chain (Q8.8 x Q8.8 fixed point), one launch per level, batch in grid.y
//   k_fast     A2  FAST-9/16 score for every pixel with DPX u16x2 min/max (VIMNMX3.U16x2), NMS, border filter,
//                  raster-ordered emission per 16-row strip (no atomics on global memory, no sort)
//   k_select1  A3  gather strips -> exact libstdc++ retainBest(2N) replay on the FAST score (one warp per
//                  (level, image), warp-parallel Hoare passes, records in shared memory)
//   k_harris   A3  Harris response of every survivor (one thread per survivor)
//   k_select2  A3  retainBest(N) replay on the Harris response
//   k_blur     A5  separable 7-tap float32 FMA Gaussian, reflect-101, round-half-even
//   k_brief    A4/A6  intensity-centroid angle + steered BRIEF (one warp per keypoint, one descriptor byte per
//                  lane) + KeyPoint assembly
//
// Everything is integer / explicitly rounded float: the file is compiled with -fmad=false and uses the
// __f*_rn intrinsics wherever the rounding sequence is part of the specification.
#pragma once
#include <cuda.h>  // CUtensorMap (the tensor maps are encoded by the host through the driver entry point, no libcuda link)
#include <cuda_runtime.h>
#include <stdint.h>

#include "select_serial.h"
#include "vslam_params.h"

#ifndef VSLAM_BLUR_MINB
#define VSLAM_BLUR_MINB 3
#endif
#ifndef VSLAM_RESIZE_MINB
#define VSLAM_RESIZE_MINB 7
#endif

namespace vk {

#define VSLAM_FULL 0xffffffffu

// rBRIEF pattern in global (not __constant__) memory: every lane reads a different test, which the constant cache would
// serialise.  TRANSPOSED for coalescing - lane b evaluates tests 8b .. 8b+7, so g_brief_pattern_i8[q][b] = test 8b + q -
// and packed as four bytes (x0, y0, x1, y1) per test (|coordinate| <= 13): one 4-byte load per test, one cache line per
// warp.  With VSLAM_BRIEF_PACKED the bytes are biased by 128 (steered_brief_byte), else two's complement.  Filled by the
// host at handle creation (upload_tables in vslam_b200.cu) from brief_pattern.inc.
__device__ unsigned g_brief_pattern_i8[8][32];

// Orientation tables (filled by the host at handle creation, see upload_tables in vslam_b200.cu):
//   g_angle_mask[o][j][r]  byte b = 0xff iff pixel u = 4j + b - (15 + o) of patch row r lies inside the 31-px disc
//                          (row index innermost: lane r reads consecutive words -> one cache line per load)
//   g_angle_u[o][j]        byte b = (int8) u
// o = (cx - 15) & 3 is the alignment of the staged patch; with them a patch row is reduced by 9 x 2 dp4a.
__device__ unsigned g_angle_mask[4][9][32];
__device__ unsigned g_angle_u[4][9];

__device__ __forceinline__ int dp4a_u8_s8(unsigned a, unsigned b, int c) {
  int d;
  asm("dp4a.u32.s32 %0, %1, %2, %3;" : "=r"(d) : "r"(a), "r"(b), "r"(c));
  return d;
}

// Non-blocking prefetch of the 32-byte sector at p into L2 (no register, no scoreboard): the register-resident
// streaming kernels below keep only 1-2 rows of loads in flight per lane, which covers the L2 latency but not a DRAM
// miss; the row that will be loaded a few iterations later is requested from DRAM here.
__device__ __forceinline__ void prefetch_l2(const void* p) { asm volatile("prefetch.global.L2 [%0];" ::"l"(p)); }

struct __align__(8) LvlKp {
  unsigned xy;  // (y<<16)|x in level coordinates
  float resp;   // Harris response
};

__device__ __forceinline__ const uint8_t* level_ptr(const PyrP& P, int l, int img, const uint8_t* in_l,
                                                    const uint8_t* in_r, const uint8_t* pyr, int& pitch) {
  if (l == 0) {
    const int f = P.eyes == 2 ? img >> 1 : img, e = P.eyes == 2 ? (img & 1) : 0;  // eyes is 1 or 2
    pitch = P.in_pitch;
    return (e ? in_r : in_l) + (long long)f * P.in_frame_stride;
  }
  pitch = P.lv[l].pitch;
  return pyr + (long long)img * P.pyr_img_stride + P.lv[l].pyr_off;
}

// ----------------------------------------------------------------------------------------------------------
// A1 — pyramid level l from level l-1 (INTER_LINEAR_EXACT: Q8.8 x Q8.8 fixed point).  coef[] holds per destination
// column / row (src offset | alpha1 << 16), computed on the host in f64.
//
// Register-resident, no shared memory, no barriers: one LANE owns 4 destination columns x (<= rs) destination rows
// and walks down.  The x coefficients, PRMT selectors and word offsets of its columns are set up once per item; the
// horizontal interpolation h(s) of a source row s is computed once (3 aligned word loads, 2 funnel shifts,
// 4 x (PRMT + DP2A)) and serves both destination rows that blend it.  The raw words of the next source row are
// prefetched while the current one is processed.
// grid = (ceil(ceil(w/4) * nstr / 128), n_img), block = 128; items are numbered column-group fastest.
// ----------------------------------------------------------------------------------------------------------
constexpr int kResizeRing = 8;  // source rows in flight per lane (cp.async ring in shared memory)
__global__ void __launch_bounds__(128, VSLAM_RESIZE_MINB) k_resize(const __grid_constant__ PyrP P, int l, int rs,
                                                const uint8_t* __restrict__ in_l, const uint8_t* __restrict__ in_r,
                                                uint8_t* __restrict__ pyr, const uint32_t* __restrict__ coef) {
  const LevelP& D = P.lv[l];
  const LevelP& S = P.lv[l - 1];
  const int img = blockIdx.y;
  const int ncg = D.pyr_ncg;
  const int item = blockIdx.x * 128 + threadIdx.x;
  const int strip = item / ncg, cg = D.pyr_cg0 + (item - strip * ncg);
  const int y0 = D.pyr_y0 + strip * rs;
  if (y0 >= D.pyr_y1) return;
  const int rows = min(rs, D.pyr_y1 - y0);
  int spitch;
  const uint8_t* src = level_ptr(P, l - 1, img, in_l, in_r, pyr, spitch);
  uint8_t* dst = pyr + (long long)img * P.pyr_img_stride + D.pyr_off + (size_t)y0 * D.pitch + 4 * cg;
  const uint32_t* yc = coef + D.ycoef_off + y0;
  const uint32_t* xc = coef + D.xcoef_off;

  // ---- per-item column setup.  The (at most 8) source bytes the four pixels need start at ox[0]: three aligned
  //      words + two funnel shifts bring them to bytes 0..7 of (v0, v1); a per-pixel PRMT selector picks
  //      (p[ox], p[ox+1]) and DP2A forms alpha0*p0 + alpha1*p1.  Word offsets are clamped to the row: a clamped word
  //      only ever supplies bytes beyond the source row, whose weight is 0.  Items whose window is wider than 8 bytes
  //      (scale factors near 2) use byte loads.
  int ox[4];
  unsigned a01[4], sel[4];
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    const unsigned c = xc[min(4 * cg + k, D.w - 1)];
    ox[k] = (int)(c & 0xffffu);
    const unsigned a1 = c >> 16;
    a01[k] = (256u - a1) | (a1 << 16);
  }
  const bool wide = ox[3] + 1 - ox[0] > 7;
  const unsigned sh = 8u * (unsigned)(ox[0] & 3);
  int wo[3];
#pragma unroll
  for (int j = 0; j < 3; ++j) wo[j] = min((ox[0] & ~3) + 4 * j, spitch - 4);
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    const unsigned e = (unsigned)(ox[k] - ox[0]) & 7u;   // byte offset of p[ox[k]] inside (v0, v1)
    sel[k] = e | (((e + 1u) & 7u) << 4) | 0x4400u;       // bytes 2,3 of the result are don't-care (DP2A.LO ignores them)
  }
  const int shm1 = S.h - 1, swm1 = S.w - 1;

  // Source rows reach the lane through a private ring in shared memory filled by cp.async (global -> shared without a
  // register): kResizeRing consecutive source rows are in flight per lane, starting at the first row the item needs.
  // Row s of the item lives in slot s % kResizeRing as one 16-byte record (3 words used), read back with one LDS.128.
  __shared__ __align__(16) uint4 s_ring[kResizeRing][128];
  const unsigned ring0 = (unsigned)__cvta_generic_to_shared(&s_ring[0][threadIdx.x]);
  const int srow0 = (int)(yc[0] & 0xffffu);       // first source row of the item
  const uint8_t* gp[3];
#pragma unroll
  for (int j = 0; j < 3; ++j) gp[j] = src + (size_t)srow0 * spitch + wo[j];
  int issued = 0;                                  // source rows enqueued so far (relative to srow0)
  auto issue_next = [&]() {
    const unsigned d = ring0 + (unsigned)(issued & (kResizeRing - 1)) * (unsigned)(128 * sizeof(uint4));
#pragma unroll
    for (int j = 0; j < 3; ++j) asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(d + 4u * j), "l"(gp[j]) : "memory");
    asm volatile("cp.async.commit_group;" ::: "memory");
    if (srow0 + issued < shm1) {  // rows past the last one re-read the last row (never consumed with weight > 0)
#pragma unroll
      for (int j = 0; j < 3; ++j) gp[j] += spitch;
    }
    ++issued;
  };
  auto hcalc = [&](int sy, unsigned (&h)[4]) {
    asm volatile("cp.async.wait_group %0;" ::"n"(kResizeRing - 1) : "memory");
    if (!wide) {
      uint4 u;
      asm volatile("ld.shared.v4.u32 {%0, %1, %2, %3}, [%4];"
                   : "=r"(u.x), "=r"(u.y), "=r"(u.z), "=r"(u.w)
                   : "r"(ring0 + (unsigned)((sy - srow0) & (kResizeRing - 1)) * (unsigned)(128 * sizeof(uint4))));
      const unsigned v0 = __funnelshift_r(u.x, u.y, sh), v1 = __funnelshift_r(u.y, u.z, sh);
#pragma unroll
      for (int k = 0; k < 4; ++k) h[k] = __dp2a_lo(a01[k], __byte_perm(v0, v1, sel[k]), 0u);
    } else {
      const uint8_t* rp = src + (size_t)sy * spitch;
#pragma unroll
      for (int k = 0; k < 4; ++k)
        h[k] = (a01[k] & 0xffffu) * rp[ox[k]] + (a01[k] >> 16) * rp[min(ox[k] + 1, swm1)];
    }
  };

#pragma unroll
  for (int i = 0; i < kResizeRing; ++i) issue_next();
  unsigned hA[4] = {0, 0, 0, 0}, hB[4] = {0, 0, 0, 0};
  int rowB = srow0 - 1;  // hB = h(rowB); the first destination row computes rows o and o+1
  unsigned yy = yc[0];
  for (int r = 0; r < rows; ++r) {
    const int o = (int)(yy & 0xffffu), o1 = min(o + 1, shm1);
    unsigned b1 = yy >> 16, b0 = 256u - b1;
    yy = yc[min(r + 1, rows - 1)];
    while (rowB < o1) {
      ++rowB;
#pragma unroll
      for (int k = 0; k < 4; ++k) hA[k] = hB[k];
      hcalc(rowB, hB);   // waits for the oldest row in flight and interpolates it horizontally
      issue_next();      // its slot is free again: enqueue the row kResizeRing ahead
    }
    if (o1 == o) {  // bottom replicate: both taps are the last source row (beta1 == 0): take it from hB
      b1 = b0;
      b0 = 0;
    }
    unsigned v[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) v[k] = b0 * hA[k] + (b1 * hB[k] + 32768u);  // < 2^24: byte 2 = result
    const unsigned word = __byte_perm(__byte_perm(v[0], v[1], 0x0062), __byte_perm(v[2], v[3], 0x0062), 0x5410);
    *reinterpret_cast<uint32_t*>(dst + (size_t)r * D.pitch) = word;
  }
  asm volatile("cp.async.wait_group 0;" ::: "memory");
}

// ----------------------------------------------------------------------------------------------------------
// A2 — FAST-9/16.  For a circle p[0..15] around centre c the FAST score is
//     S = max( max_k min_{j=k..k+8} p_j  -  c ,  c  -  min_k max_{j=k..k+8} p_j )
// (corner iff S > t, response = S-1).  min/max over the 16 circular 9-windows are computed with 3-input DPX
// min/max on u16x2 lanes: two pixels per instruction, no branches, no divergence.
// ----------------------------------------------------------------------------------------------------------
__device__ __forceinline__ unsigned max3u(unsigned a, unsigned b, unsigned c) { return __vimax3_u16x2(a, b, c); }
__device__ __forceinline__ unsigned min3u(unsigned a, unsigned b, unsigned c) { return __vimin3_u16x2(a, b, c); }

// p: 16 circle samples (u16x2 = two adjacent pixels); c2: the two centres.  Returns max(S - t, 0) per lane.
__device__ __forceinline__ unsigned fast_score2(const unsigned (&p)[16], unsigned c2, unsigned t2) {
  // The 9-windows starting at k and k+1 share the 8-window {k+1..k+8}: max(min9_k, min9_{k+1}) =
  // min(min8, max(p_k, p_{k+9})).  8-windows from pair minima at the odd positions: 36 instead of 40 DPX per polarity.
  unsigned a2[8], b2[8], a4[8], b4[8], r[8], q[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) {  // i <-> odd position j = 2i+1: {j, j+1}
    a2[i] = __vminu2(p[2 * i + 1], p[(2 * i + 2) & 15]);
    b2[i] = __vmaxu2(p[2 * i + 1], p[(2 * i + 2) & 15]);
  }
#pragma unroll
  for (int i = 0; i < 8; ++i) {  // {j .. j+3}
    a4[i] = __vminu2(a2[i], a2[(i + 1) & 7]);
    b4[i] = __vmaxu2(b2[i], b2[(i + 1) & 7]);
  }
#pragma unroll
  for (int i = 0; i < 8; ++i) {  // even k = 2i: 8-window = a4 at positions k+1 (index i) and k+5 (index i+2)
    const unsigned mx = __vmaxu2(p[2 * i], p[(2 * i + 9) & 15]), mn = __vminu2(p[2 * i], p[(2 * i + 9) & 15]);
    r[i] = min3u(a4[i], a4[(i + 2) & 7], mx);
    q[i] = max3u(b4[i], b4[(i + 2) & 7], mn);
  }
  const unsigned M9 = max3u(max3u(r[0], r[1], r[2]), max3u(r[3], r[4], r[5]), __vmaxu2(r[6], r[7]));
  const unsigned m9 = min3u(min3u(q[0], q[1], q[2]), min3u(q[3], q[4], q[5]), __vminu2(q[6], q[7]));
  unsigned bright = __vmaxu2(M9, c2) - c2;  // lane-wise non-negative: no borrow between lanes
  unsigned dark = c2 - __vminu2(m9, c2);
  unsigned S = __vmaxu2(bright, dark);
  return __vmaxu2(S, t2) - t2;
}

// 4 pixels starting at byte offset dx relative to word b, from the three consecutive words a|b|c.
template <int DX>
__device__ __forceinline__ unsigned shifted4(unsigned a, unsigned b, unsigned c) {
  if constexpr (DX < 0) return __funnelshift_r(a, b, 8u * (unsigned)(4 + DX));
  else if constexpr (DX > 0) return __funnelshift_r(b, c, 8u * (unsigned)DX);
  else return b;
}
__device__ __forceinline__ unsigned pair_lo(unsigned s) { return __byte_perm(s, 0u, 0x4140); }
__device__ __forceinline__ unsigned pair_hi(unsigned s) { return __byte_perm(s, 0u, 0x4342); }

// ---- shared-memory access by 32-bit shared-window address (plain LDS / STS with immediate offsets; generic
//      pointers to shared arrays make ptxas re-derive the window base inside every loop)
template <int OFF>
__device__ __forceinline__ unsigned lds32(unsigned a) {
  unsigned v;
  asm volatile("ld.shared.u32 %0, [%1+%2];" : "=r"(v) : "r"(a), "n"(OFF));
  return v;
}
__device__ __forceinline__ unsigned lds16(unsigned a) {
  unsigned short v;
  asm volatile("ld.shared.u16 %0, [%1];" : "=h"(v) : "r"(a));
  return v;
}
__device__ __forceinline__ unsigned lds8(unsigned a) {
  unsigned v;
  asm volatile("ld.shared.u8 %0, [%1];" : "=r"(v) : "r"(a));
  return v;
}
template <int OFF>
__device__ __forceinline__ unsigned lds16o(unsigned a) {
  unsigned short v;
  asm volatile("ld.shared.u16 %0, [%1+%2];" : "=h"(v) : "r"(a), "n"(OFF));
  return v;
}
template <int OFF>
__device__ __forceinline__ unsigned lds8o(unsigned a) {
  unsigned v;
  asm volatile("ld.shared.u8 %0, [%1+%2];" : "=r"(v) : "r"(a), "n"(OFF));
  return v;
}
__device__ __forceinline__ void sts16(unsigned a, unsigned v) {
  asm volatile("st.shared.u16 [%0], %1;" ::"r"(a), "h"((unsigned short)v));
}

// ----------------------------------------------------------------------------------------------------------
// A1, tile form — same arithmetic as k_resize, different data movement.  One BLOCK owns (rs destination rows) x (a block
// of 4-column groups) of level l: the source rows it needs (about scale * rs + 2 rows of level l-1, the column span its
// groups read) are copied global -> shared by the bulk-copy engine (cp.async.bulk, one instruction per row, completion
// on an mbarrier) and every lane then walks down its 4 columns reading 3 words per source row from shared memory with
// immediate offsets.  k_resize spends ~150 instructions per lane and destination row, most of them the 64-bit address
// arithmetic of its three per-lane cp.async per source row; here the same row costs ~40.
// grid = (strips * ncb, n_img), block = 32 * ceil(cgb / 32); dynamic smem = resize_tile_smem_bytes().
// ----------------------------------------------------------------------------------------------------------
__host__ inline size_t resize_tile_smem_bytes(int src_rows_bound, int span_bound) {
  return (size_t)src_rows_bound * (size_t)(span_bound + 16) + 32;
}
template <bool WIDE>  // WIDE: some item of the level reads more than 8 source bytes (scale factors near 2): byte loads
__global__ void __launch_bounds__(288) k_resize_tile(const __grid_constant__ PyrP P, int l, int rs, int ncb, int cgb,
                                                     const uint8_t* __restrict__ in_l, const uint8_t* __restrict__ in_r,
                                                     uint8_t* __restrict__ pyr, const uint32_t* __restrict__ coef) {
  extern __shared__ __align__(16) uint8_t smem[];
  const LevelP& D = P.lv[l];
  const LevelP& S = P.lv[l - 1];
  const int img = blockIdx.y, tid = threadIdx.x;
  const int strip = blockIdx.x / ncb, cb = blockIdx.x - strip * ncb;
  const int y0 = D.pyr_y0 + strip * rs;
  if (y0 >= D.pyr_y1) return;
  const int rows = min(rs, D.pyr_y1 - y0);
  const int cg_first = D.pyr_cg0 + cb * cgb;
  const int cg_last = min(cg_first + cgb, D.pyr_cg0 + D.pyr_ncg) - 1;  // inclusive
  if (cg_first > cg_last) return;
  int spitch;
  const uint8_t* src = level_ptr(P, l - 1, img, in_l, in_r, pyr, spitch);
  const uint32_t* yc = coef + D.ycoef_off + y0;
  const uint32_t* xc = coef + D.xcoef_off;
  int shm1 = S.h - 1, swm1 = S.w - 1;
  // source window of the block: rows [srow0, srow1], bytes [sx_lo, sx_hi) of every row (16-byte granules)
  const int srow0 = (int)(yc[0] & 0xffffu);
  const int srow1 = min((int)(yc[rows - 1] & 0xffffu) + 1, shm1);
  const int nsrc = srow1 - srow0 + 1;
  const int sx_lo = (int)(xc[4 * cg_first] & 0xffffu) & ~15;
  const int sx_hi = min(((int)(xc[min(4 * cg_last + 3, D.w - 1)] & 0xffffu) + 2 + 15) & ~15, spitch);
  const int span = sx_hi - sx_lo, SP = span + 16;
  const unsigned sbase = (unsigned)__cvta_generic_to_shared(smem) + 16u;  // [0, 8) holds the mbarrier
  const unsigned mbar = sbase - 16u;
  if (tid == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(mbar));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  if (tid < 32) {
    if (tid == 0)
      asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(mbar), "r"((unsigned)(nsrc * span)) : "memory");
    __syncwarp();
    for (int r = tid; r < nsrc; r += 32) {
      const uint8_t* g = src + (size_t)(srow0 + r) * spitch + sx_lo;
      asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                       sbase + (unsigned)(r * SP)),
                   "l"(g), "r"((unsigned)span), "r"(mbar)
                   : "memory");
    }
  }
  const int cg = cg_first + tid;
  if (cg > cg_last) return;  // no block-wide barrier below

  // ---- per-item column setup (identical to k_resize): the (at most 8) source bytes the four pixels need start at
  //      ox[0]; three words + two funnel shifts bring them to bytes 0..7 of (v0, v1), a per-pixel PRMT selector picks
  //      (p[ox], p[ox+1]) and DP2A forms alpha0*p0 + alpha1*p1.  Words that reach past the staged span only ever supply
  //      bytes whose weight is 0 (the row stride leaves 16 bytes of slack).
  int ox[4];
  unsigned a01[4], sel[4];
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    const unsigned c = xc[min(4 * cg + k, D.w - 1)];
    ox[k] = (int)(c & 0xffffu);
    const unsigned a1 = c >> 16;
    a01[k] = (256u - a1) | (a1 << 16);
  }
  unsigned sh = 8u * (unsigned)(ox[0] & 3);
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    const unsigned e = (unsigned)(ox[k] - ox[0]) & 7u;
    sel[k] = e | (((e + 1u) & 7u) << 4) | 0x4400u;
  }
  // word 0 of source row `rowB + 1` (WIDE: byte sx_lo of that row)
  unsigned arow = WIDE ? sbase - (unsigned)sx_lo : sbase + (unsigned)((ox[0] & ~3) - sx_lo);
  asm volatile("" : "+r"(sh), "+r"(arow));
  uint8_t* dst = pyr + (long long)img * P.pyr_img_stride + D.pyr_off + (size_t)y0 * D.pitch + 4 * cg;
  int dpitch = D.pitch;
  // loop invariants pinned in registers (ptxas otherwise re-derives them through indexed constant-bank loads per row)
  asm volatile("" : "+r"(shm1), "+r"(swm1), "+r"(dpitch));
  {
    unsigned done = 0;
    while (!done)
      asm volatile(
          "{\n .reg .pred p;\n mbarrier.try_wait.parity.shared::cta.b64 p, [%1], 0;\n selp.u32 %0, 1, 0, p;\n}\n"
          : "=r"(done)
          : "r"(mbar)
          : "memory");
  }
  unsigned hA[4] = {0, 0, 0, 0}, hB[4] = {0, 0, 0, 0};
  int rowB = srow0 - 1;  // hB = h(rowB)
  unsigned yy = yc[0];
  for (int r = 0; r < rows; ++r) {
    const int o = (int)(yy & 0xffffu), o1 = min(o + 1, shm1);
    unsigned b1 = yy >> 16, b0 = 256u - b1;
    yy = yc[min(r + 1, rows - 1)];
#pragma unroll 1
    while (rowB < o1) {
      ++rowB;
#pragma unroll
      for (int k = 0; k < 4; ++k) hA[k] = hB[k];
      if constexpr (!WIDE) {
        const unsigned ux = lds32<0>(arow), uy = lds32<4>(arow), uz = lds32<8>(arow);
        const unsigned v0 = __funnelshift_r(ux, uy, sh), v1 = __funnelshift_r(uy, uz, sh);
#pragma unroll
        for (int k = 0; k < 4; ++k) hB[k] = __dp2a_lo(a01[k], __byte_perm(v0, v1, sel[k]), 0u);
      } else {
#pragma unroll
        for (int k = 0; k < 4; ++k)
          hB[k] = (a01[k] & 0xffffu) * lds8(arow + (unsigned)ox[k]) + (a01[k] >> 16) * lds8(arow + (unsigned)min(ox[k] + 1, swm1));
      }
      arow += (unsigned)SP;
    }
    if (o1 == o) {  // bottom replicate: both taps are the last source row (beta1 == 0): take it from hB
      b1 = b0;
      b0 = 0;
    }
    unsigned v[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) v[k] = b0 * hA[k] + (b1 * hB[k] + 32768u);  // < 2^24: byte 2 = result
    const unsigned word = __byte_perm(__byte_perm(v[0], v[1], 0x0062), __byte_perm(v[2], v[3], 0x0062), 0x5410);
    *reinterpret_cast<uint32_t*>(dst) = word;
    dst += dpitch;
  }
}

#ifndef VSLAM_FAST_THREADS
#define VSLAM_FAST_THREADS 192  // 6 warps: one per 128-pixel column group of a 752-wide level
#endif
#ifndef VSLAM_FAST_MINB
#define VSLAM_FAST_MINB 4
#endif
#ifndef VSLAM_FAST_CL
#define VSLAM_FAST_CL 512
#endif
constexpr int kFastThreads = VSLAM_FAST_THREADS;
constexpr int kFastWarps = kFastThreads / 32;
constexpr int kFastSeg = 7;                         // score rows per phase-A segment == depth of the register window
constexpr int kFastQ = 512;                         // per-warp candidate ring (entries) >= 7 rows x 64 pairs + 31
constexpr int kFastCL = VSLAM_FAST_CL;              // per-warp pre-NMS corner list (entries)
static_assert((VSLAM_FAST_ROWS + 2) % kFastSeg == 0, "score rows of a strip must be whole segments");
static_assert(VSLAM_FAST_ROWS + 2 <= 32, "score row index is a 5-bit field of the queue entries");
static_assert(kFastQ >= kFastSeg * 64 + 31, "candidate ring too small");

// Shared-memory footprint of k_fast for a level of the given pitch.
__host__ __device__ inline int fast_sp(int pitch) { return pitch + 32; }
__host__ __device__ inline int fast_bw(int pitch) { return (pitch + 31) >> 5; }
__host__ inline size_t fast_smem_bytes(int pitch) {
  size_t sp = fast_sp(pitch), bw = fast_bw(pitch);
  return (VSLAM_FAST_ROWS + 8) * sp + (VSLAM_FAST_ROWS + 2) * sp + ((VSLAM_FAST_ROWS * bw * 4 + 15) & ~(size_t)15) +
         kFastWarps * (kFastQ + kFastCL) * 2 + 64 + (kFastWarps + 2 + VSLAM_FAST_ROWS) * 4;
}

// grid = (total_fast_strips, n_img), block = kFastThreads.  One block = one strip of VSLAM_FAST_ROWS emission rows of
// one level of one image.
//   staging   the strip's pixel rows (+-4) are one contiguous range of global memory and keep the global pitch in shared
//             memory: two cp.async.bulk copies (rows 0-13, the rest) on two mbarriers, issued before anything else; the
//             threads clear the score tile meanwhile, warp 0 polls the first barrier, one block barrier publishes it
//   phase A   every lane owns one group of 4 pixels and walks DOWN a 7-row segment with a 7-row window of raw words in
//             registers (3 LDS per step): byte-SIMD necessary condition on 4 antipodal pairs of the circle; the pixel
//             pairs that pass are appended to the warp's candidate ring (ballot compaction).  Work items = 32
//             consecutive columns of the segment-major (segment, group) list, handed out dynamically; a warp polls the
//             second staging barrier before its first item that reaches into a later segment
//   phase B   whenever the ring holds >= 32 pairs (and once at the end) 32 lanes score 32 pairs exactly (DPX);
//             non-zero scores go to the score tile and to the warp's pre-NMS corner list
//   NMS       sparse, every warp over its own corner list (full scan if a list overflowed), into a bitmask
//   emission  popcount scan of the bitmask -> raster-ordered records (== cv::FAST's output order), no atomics, no sort
__global__ void __launch_bounds__(kFastThreads, VSLAM_FAST_MINB) k_fast(const __grid_constant__ PyrP P, const uint8_t* __restrict__ in_l,
                                                       const uint8_t* __restrict__ in_r,
                                                       const uint8_t* __restrict__ pyr, int* __restrict__ strip_cnt,
                                                       uint32_t* __restrict__ strip_rec) {
  extern __shared__ __align__(16) uint8_t smem[];
  const int strip = blockIdx.x, img = blockIdx.y, tid = threadIdx.x;
  const int lane = tid & 31, warp = tid >> 5;
  int l = 0;
  for (int k = 1; k < P.nlevels && strip >= P.lv[k].fast_strip0; ++k)  // fast_strip0 is non-decreasing
    if (P.lv[k].fast_nstrips > 0) l = k;
  const LevelP& L = P.lv[l];
  const int s_local = strip - L.fast_strip0;
  const int ys = L.ey0 + s_local * VSLAM_FAST_ROWS;
  const int ye = min(ys + VSLAM_FAST_ROWS, L.ey1);
  const int rows = ye - ys;
  int pitch;
  const uint8_t* img_base = level_ptr(P, l, img, in_l, in_r, pyr, pitch);
  const int SP = fast_sp(pitch), BW = fast_bw(pitch);
  const int PP = pitch;  // pixel-tile row stride == global row stride: the tile is ONE contiguous range of global memory
  // layout (byte offsets into the dynamic shared memory)
  const int o_sc = (VSLAM_FAST_ROWS + 8) * SP;               // pixel tile: image row ys-4+r at r*PP + 16
  const int o_bits = o_sc + (VSLAM_FAST_ROWS + 2) * SP;      // score tile: image row ys-1+r at o_sc + r*SP + 16
  const int o_q = o_bits + ((VSLAM_FAST_ROWS * BW * 4 + 15) & ~15);  // corner bitmask; then the per-warp candidate rings
  const int o_cl = o_q + kFastWarps * kFastQ * 2;            // per-warp corner lists
  const int o_misc = (o_cl + kFastWarps * kFastCL * 2 + 15) & ~15;  // mbarrier (8 B), list counts, total
  const unsigned sbase = (unsigned)__cvta_generic_to_shared(smem);
  uint8_t* sc = smem + o_sc;
  uint32_t* bits = reinterpret_cast<uint32_t*>(smem + o_bits);
  int* s_ccnt = reinterpret_cast<int*>(smem + o_misc + 16);  // [kFastWarps] corner-list lengths
  int* s_next = s_ccnt + kFastWarps;
  int* s_rowcnt = s_next + 1;                                // [VSLAM_FAST_ROWS] NMS survivors per emission row                               // dynamic work-item counter of phase A
  const unsigned mbar = sbase + (unsigned)o_misc;  // two mbarriers: tile rows [0, n0) and [n0, rows + 8)

  // ---- staging: bulk copies of the rows ys-4 .. ye+3 (always inside the image: ey0 >= 4, ey1 <= h - 4).  Two copies:
  // the first 14 rows (all that segment 0 of phase A reads) complete on mbarrier 0, the rest on mbarrier 1.  Thread 0
  // initialises both barriers and issues both copies at once; warp 0 is the only poller of barrier 0, the block barrier
  // behind its poll publishes the rows (and the cleared score tile, and the barriers' initialisation) to the other warps.
  const int n0 = min(2 * kFastSeg, rows + 8), n1 = rows + 8 - n0;
  if (tid == 0) {
    *s_next = 0;
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(mbar));
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(mbar + 8u));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    const uint8_t* g = img_base + (size_t)(ys - 4) * pitch;
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(mbar), "r"((unsigned)(n0 * pitch)) : "memory");
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(sbase + 16u),
                 "l"(g), "r"((unsigned)(n0 * pitch)), "r"(mbar)
                 : "memory");
    if (n1 > 0) {
      asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(mbar + 8u), "r"((unsigned)(n1 * pitch))
                   : "memory");
      asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                       sbase + 16u + (unsigned)(n0 * pitch)),
                   "l"(g + (size_t)n0 * pitch), "r"((unsigned)(n1 * pitch)), "r"(mbar + 8u)
                   : "memory");
    }
  }
  for (int i = tid; i < ((VSLAM_FAST_ROWS + 2) * SP) >> 4; i += kFastThreads)
    reinterpret_cast<uint4*>(sc)[i] = make_uint4(0, 0, 0, 0);
  for (int i = tid; i < (VSLAM_FAST_ROWS * BW + 3) >> 2; i += kFastThreads)  // 128-bit stores; the region is padded
    reinterpret_cast<uint4*>(bits)[i] = make_uint4(0, 0, 0, 0);
  if (tid < VSLAM_FAST_ROWS) s_rowcnt[tid] = 0;
  auto wait_rows = [&](unsigned bar) {  // acquire the rows whose copy completes on this mbarrier
    unsigned done = 0;
    while (!done)
      asm volatile(
          "{\n .reg .pred p;\n mbarrier.try_wait.parity.shared::cta.b64 p, [%1], 0;\n selp.u32 %0, 1, 0, p;\n}\n"
          : "=r"(done)
          : "r"(bar)
          : "memory");
  };
  if (warp == 0) {
    __syncwarp();  // lanes 1..31 must not look at the barrier before lane 0 has initialised it
    wait_rows(mbar);
  }
  __syncthreads();
  // the other warps observe the (by now completed) phase themselves: one try_wait that succeeds at once, so that every
  // thread that reads the bulk-copied rows has acquired them through the mbarrier, not only through the block barrier
  if (warp != 0) wait_rows(mbar);

  // ---- scores on the emission region dilated by one pixel (always inside FAST's own 3-px border because ex0 >= 4).
  // Phase A: a 9-arc of the 16-circle contains at least one point of every antipodal pair {k, k+8}, so
  //   (|d_k| > t or |d_{k+8}| > t) must hold for each pair; the four pairs k = 0, 2, 4, 6 are tested (22 % of the pixel
  //   pairs of the synthetic scenes survive, 10 % contain a corner).  Pixels that fail keep score 0 — exactly what the
  //   full computation would give them.
  const int sx0 = max(3, L.ex0 - 1), sx1 = min(L.w - 3, L.ex1 + 1);
  const int sy_lo = max(3, ys - 1), sy_hi = min(L.h - 3, ye + 1);  // [sy_lo, sy_hi)
  const int gx0 = sx0 >> 2, ngx = ((sx1 - 1) >> 2) - gx0 + 1;
  const int nsr = sy_hi - sy_lo;                                   // <= VSLAM_FAST_ROWS + 2
  const unsigned t = (unsigned)P.fast_threshold;
  const unsigned t2 = t * 0x00010001u;
  const unsigned tadd = (127u - t) * 0x01010101u;  // fast_threshold <= 127 (checked by vslam_create)
  const unsigned lt = (1u << lane) - 1u;
  const unsigned pix0 = sbase + (unsigned)((sy_lo - (ys - 4)) * PP + 16);  // centre pixel of score row 0, x = 0
  const unsigned sc0 = sbase + (unsigned)(o_sc + (sy_lo - (ys - 1)) * SP + 16);
  const unsigned qb = sbase + (unsigned)(o_q + warp * kFastQ * 2);
  const unsigned clb = sbase + (unsigned)(o_cl + warp * kFastCL * 2);
  int qn = 0, qhead = 0, cn = 0;  // warp-uniform: ring fill / head, corner-list length

  // exact scores of up to 32 queued pairs, one pair per lane
  auto phase_b = [&]() {
    unsigned flag = 0, ccode = 0;
    if (lane < qn) {
      const unsigned code = lds16(qb + 2u * (unsigned)((qhead + lane) & (kFastQ - 1)));
      const unsigned half = code & 1u, g4 = (code << 1) & 0xffcu, ry = code >> 11;  // g4 = 4 * group
      const unsigned rowo = ry * (unsigned)PP + g4;
      // byte address of the pair's first pixel; row bases y-3 .. y+3.  Even dx: one aligned 16-bit load + PRMT; odd dx:
      // two byte loads combined by an IMAD (FMA pipe) - the ALU pipe, which bounds this kernel, only sees 7 PRMT.
      const unsigned A = pix0 + rowo + 2u * half;
      unsigned rb[7];
#pragma unroll
      for (int j = 0; j < 7; ++j) rb[j] = A + (unsigned)(j - 3) * (unsigned)PP;
      unsigned p[16];
#define VS_G(k, dx, dy)                                                                  \
  if constexpr ((dx) & 1) {                                                              \
    const unsigned lo = lds8o<(dx)>(rb[(dy) + 3]), hi = lds8o<(dx) + 1>(rb[(dy) + 3]);   \
    asm("mad.lo.u32 %0, %1, 65536, %2;" : "=r"(p[k]) : "r"(hi), "r"(lo));                \
  } else {                                                                               \
    p[k] = pair_lo(lds16o<(dx)>(rb[(dy) + 3]));                                          \
  }
      VS_G(0, 0, 3) VS_G(1, 1, 3) VS_G(2, 2, 2) VS_G(3, 3, 1) VS_G(4, 3, 0) VS_G(5, 3, -1) VS_G(6, 2, -2) VS_G(7, 1, -3)
      VS_G(8, 0, -3) VS_G(9, -1, -3) VS_G(10, -2, -2) VS_G(11, -3, -1) VS_G(12, -3, 0) VS_G(13, -3, 1) VS_G(14, -2, 2)
      VS_G(15, -1, 3)
#undef VS_G
      const unsigned c2 = pair_lo(lds16(A));
      const unsigned xs = fast_score2(p, c2, t2);
      const unsigned out16 = __byte_perm(xs, 0u, 0x4420);                     // (lane0, lane1) low bytes -> one u16
      const unsigned xo = g4 + 2u * half;
      sts16(sc0 + ry * (unsigned)SP + xo, out16);
      // corner-list entry: (score row << 11) | x; rows / columns outside the emission region are dropped by the NMS
      ccode = (ry << 11) | xo;
      flag = ((out16 & 0x00ffu) ? 1u : 0u) | ((out16 & 0xff00u) ? 2u : 0u);
    }
    const unsigned m0 = __ballot_sync(VSLAM_FULL, flag & 1u), m1 = __ballot_sync(VSLAM_FULL, flag & 2u);
    const int n0 = __popc(m0), pos0 = cn + __popc(m0 & lt), pos1 = cn + n0 + __popc(m1 & lt);
    if ((flag & 1u) && pos0 < kFastCL) sts16(clb + 2u * (unsigned)pos0, ccode);
    if ((flag & 2u) && pos1 < kFastCL) sts16(clb + 2u * (unsigned)pos1, ccode + 1u);
    cn += n0 + __popc(m1);
    const int done = min(qn, 32);
    qhead = (qhead + done) & (kFastQ - 1);
    qn -= done;
    __syncwarp();
  };

  // Work items = (32-group column block, 7-row segment), handed out dynamically (one shared atomic per item) so that
  // the warps of a block finish together whatever the level width; every item refills its register window (6 rows).
  // items = 32 consecutive (segment, group) columns of the segment-major list: no idle lanes in the last column block
  // of a level whose width is not a multiple of 128 pixels (10 % of the lane slots on the 8-level 752-wide pyramid)
  const int nseg = (nsr + kFastSeg - 1) / kFastSeg, nlin = nseg * ngx, nitems = (nlin + 31) >> 5;
  const unsigned inv_ngx = 0xffffffffu / (unsigned)ngx + 1u;  // ceil(2^32 / ngx): exact quotients for li < 2^16
  bool got1 = n1 == 0;  // warp-uniform: this warp has seen the second staging barrier complete
  while (true) {
    int item = 0;
    if (lane == 0) item = atomicAdd(s_next, 1);
    item = __shfl_sync(VSLAM_FULL, item, 0);
    if (item >= nitems) break;
    const int li = min(item * 32 + lane, nlin - 1);
    const bool has = item * 32 + lane < nlin;
    const int seg = (int)__umulhi((unsigned)li, inv_ngx), gi = li - seg * ngx, r0 = seg * kFastSeg;
    const int g = gx0 + gi;
    const int seg_hi = (int)__umulhi((unsigned)min(item * 32 + 31, nlin - 1), inv_ngx);
    if (!got1 && seg_hi > 0) {  // segments >= 1 read tile rows >= 14
      wait_rows(mbar + 8u);
      got1 = true;
    }
    // per-byte mask of the scored columns [sx0, sx1) of this group (only the first / last group is partial)
    unsigned colmask = 0;
#pragma unroll
    for (int b = 0; b < 4; ++b)
      if (has && 4 * g + b >= sx0 && 4 * g + b < sx1) colmask |= 0x80u << (8 * b);
    const unsigned codeg = (unsigned)g << 1;
    unsigned arow = pix0 + 4u * (unsigned)g + (unsigned)((r0 - 3) * PP);  // word b of pixel row (sy_lo + r0 - 3)
    unsigned W[7][3];  // raw words (x-4, x, x+4) of the pixel rows y-3 .. y+3; every index is a compile-time constant
#pragma unroll
    for (int j = 0; j < 6; ++j) {
      W[j][0] = lds32<-4>(arow);
      W[j][1] = lds32<0>(arow);
      W[j][2] = lds32<4>(arow);
      arow += (unsigned)PP;
    }
#pragma unroll
    for (int ph = 0; ph < kFastSeg; ++ph) {
      // window slot of pixel row y-3+j is (ph + j) % 7
#define WR(j, c) W[((ph) + (j)) % 7][c]
      WR(6, 0) = lds32<-4>(arow);
      WR(6, 1) = lds32<0>(arow);
      WR(6, 2) = lds32<4>(arow);
      arow += (unsigned)PP;
      const unsigned cw = WR(3, 1);
      const unsigned a0 = __vabsdiffu4(WR(6, 1), cw), a8 = __vabsdiffu4(WR(0, 1), cw);
      const unsigned a4 = __vabsdiffu4(shifted4<3>(WR(3, 0), WR(3, 1), WR(3, 2)), cw);
      const unsigned a12 = __vabsdiffu4(shifted4<-3>(WR(3, 0), WR(3, 1), WR(3, 2)), cw);
      const unsigned a2 = __vabsdiffu4(shifted4<2>(WR(5, 0), WR(5, 1), WR(5, 2)), cw);
      const unsigned a14 = __vabsdiffu4(shifted4<-2>(WR(5, 0), WR(5, 1), WR(5, 2)), cw);
      const unsigned a6 = __vabsdiffu4(shifted4<2>(WR(1, 0), WR(1, 1), WR(1, 2)), cw);
      const unsigned a10 = __vabsdiffu4(shifted4<-2>(WR(1, 0), WR(1, 1), WR(1, 2)), cw);
#undef WR
      // per byte "v > t" (t <= 127) as bit 7 of (v + 127 - t) | v.  The add is not masked per byte: a carry out of
      // a byte with v >= 129 + t can only SET the flag of its neighbour (a false positive of a necessary test, i.e.
      // one superfluous exact evaluation), and that byte's own flag is still set through "| v" (v >= 128).
      const unsigned f08 = (a0 + tadd) | a0 | (a8 + tadd) | a8;
      const unsigned f412 = (a4 + tadd) | a4 | (a12 + tadd) | a12;
      const unsigned f210 = (a2 + tadd) | a2 | (a10 + tadd) | a10;
      const unsigned f614 = (a6 + tadd) | a6 | (a14 + tadd) | a14;
      const int ry = r0 + ph;
      const unsigned pass = (ry < nsr) ? (f08 & f412 & f210 & f614 & colmask) : 0u;
      const unsigned mlo = __ballot_sync(VSLAM_FULL, (pass & 0x00008080u) != 0u);
      const unsigned mhi = __ballot_sync(VSLAM_FULL, (pass & 0x80800000u) != 0u);
      const unsigned code = ((unsigned)ry << 11) | codeg;
      const int tail = qhead + qn, nlo = __popc(mlo);
      if (mlo & (1u << lane)) sts16(qb + 2u * (unsigned)((tail + __popc(mlo & lt)) & (kFastQ - 1)), code);
      if (mhi & (1u << lane)) sts16(qb + 2u * (unsigned)((tail + nlo + __popc(mhi & lt)) & (kFastQ - 1)), code | 1u);
      qn += nlo + __popc(mhi);
    }
    __syncwarp();
    while (qn >= 32) phase_b();
  }
  while (qn > 0) phase_b();
  if (lane == 0) s_ccnt[warp] = cn;
  __syncthreads();

  // ---- non-maximum suppression (strictly greater than the 8 neighbours) + runByImageBorder
  bool sparse = L.w <= 2048;  // x is an 11-bit field of the corner-list entries
#pragma unroll
  for (int w = 0; w < kFastWarps; ++w) sparse = sparse && s_ccnt[w] <= kFastCL;
  if (sparse) {
    const int ry_lo = ys - sy_lo, ry_hi = ye - sy_lo;  // score rows that are emission rows: [ry_lo, ry_hi)
    // one flat index space over the warps' lists; the eight neighbours are loaded unconditionally (one shared-memory
    // latency instead of a short-circuit chain of up to eight)
    // every warp walks its OWN corner list (cn is still in a register): no list selection, no count exchange
    for (int off = lane; off < cn; off += 32) {
      const int w = warp;
      const unsigned code = lds16(sbase + (unsigned)(o_cl + (w * kFastCL + off) * 2));
      const int ry = (int)(code >> 11), x = (int)(code & 0x7ffu);
      if (ry < ry_lo || ry >= ry_hi || x < L.ex0 || x >= L.ex1) continue;
      const int er = ry - ry_lo;  // emission row
      const unsigned ca = sbase + (unsigned)(o_sc + (er + 1) * SP + 16 + x);
      const unsigned up = ca - (unsigned)SP, dn = ca + (unsigned)SP;
      const unsigned v = lds8o<0>(ca);
      const unsigned n0 = lds8o<-1>(ca), n1 = lds8o<1>(ca), n2 = lds8o<-1>(up), n3 = lds8o<0>(up), n4 = lds8o<1>(up),
                     n5 = lds8o<-1>(dn), n6 = lds8o<0>(dn), n7 = lds8o<1>(dn);
      const unsigned m = max(max(max(n0, n1), max(n2, n3)), max(max(n4, n5), max(n6, n7)));
      if (v > m) {
        atomicOr(&bits[er * BW + (x >> 5)], 1u << (x & 31));
        atomicAdd(&s_rowcnt[er], 1);
      }
    }
  } else {
    const int egx0 = L.ex0 >> 2, negx = ((L.ex1 - 1) >> 2) - egx0 + 1;
    const unsigned inv_negx = (unsigned)((0x100000000ull + (unsigned)negx - 1u) / (unsigned)negx);
    for (int it = tid; it < rows * negx; it += kFastThreads) {
      const int ry = (int)__umulhi((unsigned)it, inv_negx);
      const int g = egx0 + (it - ry * negx);
      const uint8_t* srow = sc + (ry + 1) * SP + 16;
      unsigned word = *reinterpret_cast<const uint32_t*>(srow + 4 * g);
      if (word == 0) continue;
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        int v = (word >> (8 * k)) & 0xff;
        int x = 4 * g + k;
        if (v == 0 || x < L.ex0 || x >= L.ex1) continue;
        const uint8_t* c = srow + x;
        bool keep = v > c[-1] && v > c[1] && v > c[-SP - 1] && v > c[-SP] && v > c[-SP + 1] && v > c[SP - 1] &&
                    v > c[SP] && v > c[SP + 1];
        if (keep) {
          atomicOr(&bits[ry * BW + (x >> 5)], 1u << (x & 31));
          atomicAdd(&s_rowcnt[ry], 1);
        }
      }
    }
  }
  __syncthreads();

  // ---- raster-ordered emission.  Order inside the strip == (y, x) == cv::FAST's output order.  The NMS pass counted
  //      the survivors of every emission row (s_rowcnt); every warp scans those <= 32 counts (5 shuffles, redundantly:
  //      no single-warp phase, no extra barrier) and then emits the rows r = warp, warp + nwarps, ...: lane j takes
  //      word j of the row, a warp prefix sum of the popcounts gives its offset.  (POPC issues once per 8 clk per
  //      scheduler on B200: one per word here, not one per word per warp.)
  {
    const int rc = lane < rows ? s_rowcnt[lane] : 0;
    int incl = rc;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
      const int o = __shfl_up_sync(VSLAM_FULL, incl, d);
      if (lane >= d) incl += o;
    }
    const int excl = incl - rc;  // lane r: records of the rows above row r
    uint32_t* out = strip_rec + (long long)img * P.rec_img_stride + L.rec_off + (long long)s_local * L.strip_cap;
    if (BW <= 32) {
      // the warp's rows (er = warp + k * kFastWarps) side by side: their prefix scans and their record loops interleave,
      // so the shuffle and shared-load latencies of one row hide behind the other rows
      constexpr int RPW = (VSLAM_FAST_ROWS + kFastWarps - 1) / kFastWarps;
      unsigned bwv[RPW];
      int pin[RPW], off[RPW];
#pragma unroll
      for (int k = 0; k < RPW; ++k) {
        const int er = warp + k * kFastWarps;
        bwv[k] = (er < rows && lane < BW) ? bits[er * BW + lane] : 0u;
        pin[k] = __popc(bwv[k]);
        off[k] = -pin[k];
      }
#pragma unroll
      for (int d = 1; d < 32; d <<= 1) {
#pragma unroll
        for (int k = 0; k < RPW; ++k) {
          const int o = __shfl_up_sync(VSLAM_FULL, pin[k], d);
          if (lane >= d) pin[k] += o;
        }
      }
#pragma unroll
      for (int k = 0; k < RPW; ++k) off[k] += pin[k] + __shfl_sync(VSLAM_FULL, excl, min(warp + k * kFastWarps, 31));
      const unsigned srow0 = sbase + (unsigned)(o_sc + (warp + 1) * SP + 16 + (lane << 5));
      const unsigned tm1 = (unsigned)P.fast_threshold - 1u;
      unsigned any = 1u;
      while (any) {
        // branch-free body: the RPW score loads are issued back to back (an exhausted row reads the pad byte in front
        // of its word), only the stores are predicated
        unsigned scv[RPW];
        int bp[RPW];
#pragma unroll
        for (int k = 0; k < RPW; ++k) {
          bp[k] = __ffs(bwv[k]) - 1;
          scv[k] = lds8(srow0 + (unsigned)(k * kFastWarps * SP + bp[k]));
        }
        any = 0u;
#pragma unroll
        for (int k = 0; k < RPW; ++k) {
          const bool act = bwv[k] != 0u;
          const unsigned rec = ((scv[k] + tm1) << 24) | ((unsigned)(ys + warp + k * kFastWarps) << 12) |
                               (unsigned)((lane << 5) + bp[k]);  // score S - 1 <= 254
          if (act) out[off[k]] = rec;
          off[k] += act ? 1 : 0;
          bwv[k] &= bwv[k] - 1u;
          any |= bwv[k];
        }
      }
    } else
    for (int er = warp; er < rows; er += kFastWarps) {
      int run = __shfl_sync(VSLAM_FULL, excl, er);
      for (int j0 = 0; j0 < BW; j0 += 32) {
        const int j = j0 + lane;
        unsigned bw = j < BW ? bits[er * BW + j] : 0u;
        const int c = __popc(bw);
        int pin = c;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
          const int o = __shfl_up_sync(VSLAM_FULL, pin, d);
          if (lane >= d) pin += o;
        }
        int off = run + pin - c;
        const uint8_t* srow = sc + (er + 1) * SP + 16 + (j << 5);
        while (bw) {
          const int bpos = __ffs(bw) - 1;
          bw &= bw - 1;
          const unsigned score = (unsigned)srow[bpos] + (unsigned)P.fast_threshold - 1u;  // S - 1 <= 254
          out[off++] = (score << 24) | ((unsigned)(ys + er) << 12) | (unsigned)((j << 5) + bpos);
        }
        run += __shfl_sync(VSLAM_FULL, pin, 31);
      }
    }
    if (tid == 31) strip_cnt[(long long)img * P.total_fast_strips + strip] = incl;
  }
}

// ----------------------------------------------------------------------------------------------------------
// A3 — exact replay of KeyPointsFilter::retainBest (libstdc++ std::nth_element + std::partition) by ONE warp.
// A Hoare pass of __unguarded_partition pairs the k-th "left stopper" (v[i] <= pivot, ascending) with the k-th
// "right stopper" (v[j] >= pivot, descending) for as long as L_k < R_k; every swap touches a distinct pair, so the
// pass is: two ordered stream compactions (ballot + popc), a count, and a parallel swap.  The cut the serial
// algorithm returns is min(L_{K+1}, R_K).  (Checked against libstdc++ in tests/test_select_replay.py.)
// ----------------------------------------------------------------------------------------------------------
template <typename T, typename IdxT>
__device__ int warp_partition_pivot(T* rec, int first, int last, IdxT* Ll, IdxT* Rl, int lane) {
  if (lane == 0) vsel::median_to_first(rec, first, first + 1, first + (last - first) / 2, last - 1);
  __syncwarp();
  const T pivot = rec[first];
  const unsigned lt = (1u << lane) - 1u;
  int nR = 0, nL = 0;
  // right stoppers: !(pivot > v[j]), j descending over [first, last);  left stoppers: !(v[i] > pivot), i ascending
  // over [first+1, last).  Both scans advance in the same iteration so their loads overlap.
  const int len = last - first;
  for (int o = 0; o < len; o += 32) {
    const int j = last - 1 - o - lane;
    const int i = first + 1 + o + lane;
    const bool inj = j >= first, ini = i < last;
    const T vj = inj ? rec[j] : pivot, vi = ini ? rec[i] : pivot;
    const bool sr = inj && !vsel::comp(pivot, vj);
    const bool sl = ini && !vsel::comp(vi, pivot);
    const unsigned mr = __ballot_sync(VSLAM_FULL, sr), ml = __ballot_sync(VSLAM_FULL, sl);
    if (sr) Rl[nR + __popc(mr & lt)] = (IdxT)j;
    if (sl) Ll[nL + __popc(ml & lt)] = (IdxT)i;
    nR += __popc(mr);
    nL += __popc(ml);
  }
  __syncwarp();
  const int nmin = min(nL, nR);
  int K = 0;
  for (int k0 = 0; k0 < nmin; k0 += 32) {  // L ascending, R descending: the predicate is a prefix
    int k = k0 + lane;
    bool ok = (k < nmin) && ((int)Ll[k] < (int)Rl[k]);
    unsigned m = __ballot_sync(VSLAM_FULL, ok);
    K += __popc(m);
    if (m != VSLAM_FULL) break;
  }
  int cut = last;
  if (K < nL) cut = min(cut, (int)Ll[K]);
  if (K > 0) cut = min(cut, (int)Rl[K - 1]);
  for (int k = lane; k < K; k += 32) {
    int a = (int)Ll[k], b = (int)Rl[k];
    T t = rec[a];
    rec[a] = rec[b];
    rec[b] = t;
  }
  __syncwarp();
  return cut;
}

// Returns the retained count; rec[0..count) is exactly what std::vector<KeyPoint> holds after retainBest(n_points).
template <typename T, typename IdxT>
__device__ int warp_retain_best(T* rec, int n, int n_points, IdxT* Ll, IdxT* Rl, int lane) {
  if (n_points < 0 || n <= n_points) return n;
  if (n_points == 0) return 0;
  const int nth = n_points - 1;
  int first = 0, last = n;
  int depth = 2 * vsel::floor_log2(n);
  bool sorted_tail = true;
  while (last - first > 3) {
    if (depth == 0) {  // std::__heap_select fallback (serial; practically never taken)
      if (lane == 0) vsel::heap_select_fallback(rec, first, nth, last);
      sorted_tail = false;
      break;
    }
    --depth;
    int cut = warp_partition_pivot<T, IdxT>(rec, first, last, Ll, Rl, lane);
    if (cut <= nth) first = cut;
    else last = cut;
  }
  if (sorted_tail && lane == 0) vsel::insertion_sort(rec, first, last);
  __syncwarp();
  // std::partition(v + n_points, end, resp >= thr): the k-th failing element from the left swaps with the k-th
  // passing element from the right while A_k < B_k; only the kept front part is materialised.
  const T thr = rec[nth];
  const unsigned lt = (1u << lane) - 1u;
  int nB = 0, nA = 0;
  for (int o = 0; o < n - n_points; o += 32) {
    const int j = n - 1 - o - lane, i = n_points + o + lane;
    const bool inj = j >= n_points, ini = i < n;
    const bool sb = inj && !vsel::comp(thr, rec[inj ? j : nth]);   // resp >= thr
    const bool sa = ini && vsel::comp(thr, rec[ini ? i : nth]);    // !(resp >= thr)
    const unsigned mb = __ballot_sync(VSLAM_FULL, sb), ma = __ballot_sync(VSLAM_FULL, sa);
    if (sb) Rl[nB + __popc(mb & lt)] = (IdxT)j;
    if (sa) Ll[nA + __popc(ma & lt)] = (IdxT)i;
    nB += __popc(mb);
    nA += __popc(ma);
  }
  __syncwarp();
  const int nmin = min(nA, nB);
  for (int k = lane; k < nmin; k += 32) {
    int a = (int)Ll[k], b = (int)Rl[k];
    if (a < b) rec[a] = rec[b];
  }
  __syncwarp();
  return n_points + nB;
}

// ---- the same replay by a whole BLOCK of 128 threads (4 warps).  The stopper lists are ordered by scan position, so
//      warp w builds the part of both lists that falls into its quarter of the positions: a counting pass, a 4-entry
//      prefix over the warps, then the writing pass; K is found by every thread with the same binary search (the
//      predicate L[k] < R[k] is a prefix), the swaps are spread over the block.  Ranges shorter than 128 fall back to
//      the single-warp pass (warp 0; the other warps wait at the barrier that publishes the cut).
//      All 128 threads must call these functions with identical arguments.  s_w: shared int[12].
template <typename T, typename IdxT>
__device__ int block_partition_pivot(T* rec, int first, int last, IdxT* Ll, IdxT* Rl, int* s_w) {
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int len = last - first;
  if (len < 128) {
    if (warp == 0) {
      const int cut = warp_partition_pivot<T, IdxT>(rec, first, last, Ll, Rl, lane);
      if (lane == 0) s_w[8] = cut;
    }
    __syncthreads();
    const int cut = s_w[8];
    __syncthreads();  // s_w[8] is rewritten by the next pass
    return cut;
  }
  if (tid == 0) vsel::median_to_first(rec, first, first + 1, first + len / 2, last - 1);
  __syncthreads();
  const T pivot = rec[first];
  const unsigned lt = (1u << lane) - 1u;
  const int chunk = ((len + 127) >> 7) << 5;  // scan positions per warp (multiple of 32)
  const int o_beg = warp * chunk, o_end = min(o_beg + chunk, len);
  int nR = 0, nL = 0;
  for (int o = o_beg; o < o_end; o += 32) {
    const int j = last - 1 - o - lane, i = first + 1 + o + lane;
    const bool inj = j >= first, ini = i < last;  // (o + lane < len <=> j >= first; i < last is the tighter bound for L)
    const T vj = inj ? rec[j] : pivot, vi = ini ? rec[i] : pivot;
    nR += __popc(__ballot_sync(VSLAM_FULL, inj && !vsel::comp(pivot, vj)));
    nL += __popc(__ballot_sync(VSLAM_FULL, ini && !vsel::comp(vi, pivot)));
  }
  if (lane == 0) {
    s_w[warp] = nR;
    s_w[4 + warp] = nL;
  }
  __syncthreads();
  int baseR = 0, baseL = 0, totR = 0, totL = 0;
#pragma unroll
  for (int w = 0; w < 4; ++w) {
    const int r = s_w[w], l = s_w[4 + w];
    if (w < warp) {
      baseR += r;
      baseL += l;
    }
    totR += r;
    totL += l;
  }
  nR = baseR;
  nL = baseL;
  for (int o = o_beg; o < o_end; o += 32) {
    const int j = last - 1 - o - lane, i = first + 1 + o + lane;
    const bool inj = j >= first, ini = i < last;
    const T vj = inj ? rec[j] : pivot, vi = ini ? rec[i] : pivot;
    const bool sr = inj && !vsel::comp(pivot, vj);
    const bool sl = ini && !vsel::comp(vi, pivot);
    const unsigned mr = __ballot_sync(VSLAM_FULL, sr), ml = __ballot_sync(VSLAM_FULL, sl);
    if (sr) Rl[nR + __popc(mr & lt)] = (IdxT)j;
    if (sl) Ll[nL + __popc(ml & lt)] = (IdxT)i;
    nR += __popc(mr);
    nL += __popc(ml);
  }
  __syncthreads();
  const int nmin = min(totL, totR);
  int lo = 0, hi = nmin;  // smallest k with !(L[k] < R[k]); every thread runs the same search
  while (lo < hi) {
    const int mid = (lo + hi) >> 1;
    if ((int)Ll[mid] < (int)Rl[mid]) lo = mid + 1;
    else hi = mid;
  }
  const int K = lo;
  int cut = last;
  if (K < totL) cut = min(cut, (int)Ll[K]);
  if (K > 0) cut = min(cut, (int)Rl[K - 1]);
  for (int k = tid; k < K; k += 128) {
    const int a = (int)Ll[k], b = (int)Rl[k];
    const T t = rec[a];
    rec[a] = rec[b];
    rec[b] = t;
  }
  __syncthreads();
  return cut;
}

template <typename T, typename IdxT>
__device__ int block_retain_best(T* rec, int n, int n_points, IdxT* Ll, IdxT* Rl, int* s_w) {
  if (n_points < 0 || n <= n_points) return n;
  if (n_points == 0) return 0;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int nth = n_points - 1;
  int first = 0, last = n;
  int depth = 2 * vsel::floor_log2(n);
  bool sorted_tail = true;
  while (last - first > 3) {
    if (depth == 0) {  // std::__heap_select fallback (serial; practically never taken)
      if (tid == 0) vsel::heap_select_fallback(rec, first, nth, last);
      sorted_tail = false;
      break;
    }
    --depth;
    const int cut = block_partition_pivot<T, IdxT>(rec, first, last, Ll, Rl, s_w);
    if (cut <= nth) first = cut;
    else last = cut;
  }
  if (sorted_tail && tid == 0) vsel::insertion_sort(rec, first, last);
  __syncthreads();
  // std::partition(v + n_points, end, resp >= thr), block-parallel like the Hoare pass
  const T thr = rec[nth];
  const unsigned lt = (1u << lane) - 1u;
  const int len = n - n_points;
  const int chunk = ((len + 127) >> 7) << 5;
  const int o_beg = warp * chunk, o_end = min(o_beg + chunk, len);
  int nB = 0, nA = 0;
  for (int o = o_beg; o < o_end; o += 32) {
    const int j = n - 1 - o - lane, i = n_points + o + lane;
    const bool inj = j >= n_points, ini = i < n;
    nB += __popc(__ballot_sync(VSLAM_FULL, inj && !vsel::comp(thr, rec[inj ? j : nth])));
    nA += __popc(__ballot_sync(VSLAM_FULL, ini && vsel::comp(thr, rec[ini ? i : nth])));
  }
  if (lane == 0) {
    s_w[warp] = nB;
    s_w[4 + warp] = nA;
  }
  __syncthreads();
  int baseB = 0, baseA = 0, totB = 0, totA = 0;
#pragma unroll
  for (int w = 0; w < 4; ++w) {
    const int b = s_w[w], a = s_w[4 + w];
    if (w < warp) {
      baseB += b;
      baseA += a;
    }
    totB += b;
    totA += a;
  }
  nB = baseB;
  nA = baseA;
  for (int o = o_beg; o < o_end; o += 32) {
    const int j = n - 1 - o - lane, i = n_points + o + lane;
    const bool inj = j >= n_points, ini = i < n;
    const bool sb = inj && !vsel::comp(thr, rec[inj ? j : nth]);
    const bool sa = ini && vsel::comp(thr, rec[ini ? i : nth]);
    const unsigned mb = __ballot_sync(VSLAM_FULL, sb), ma = __ballot_sync(VSLAM_FULL, sa);
    if (sb) Rl[nB + __popc(mb & lt)] = (IdxT)j;
    if (sa) Ll[nA + __popc(ma & lt)] = (IdxT)i;
    nB += __popc(mb);
    nA += __popc(ma);
  }
  __syncthreads();
  const int nmin = min(totA, totB);
  for (int k = tid; k < nmin; k += 128) {
    const int a = (int)Ll[k], b = (int)Rl[k];
    if (a < b) rec[a] = rec[b];
  }
  __syncthreads();
  return n_points + totB;
}

// A3(iii) — HarrisResponses (blockSize 7, k = 0.04) on the un-blurred level; float formula without FMA.
__device__ float harris_response(const uint8_t* __restrict__ img, int pitch, int x, int y, float scale4) {
  // Ix = [1 2 1]^T (rows) applied to dx[r][j] = v[r][j+1] - v[r][j-1];  Iy = sm[r+1][j] - sm[r-1][j] with
  // sm[r][j] = v[r][j-1] + 2 v[r][j] + v[r][j+1].  Three rows of (dx, sm) are kept live while walking 9 rows.
  // the 9 bytes x-4 .. x+4 of a row come from three aligned words (one L1 wavefront each instead of nine byte loads:
  // the threads of a warp sit on different corners, so every load instruction touches 32 different lines)
  const uint8_t* p0 = img + (size_t)(y - 4) * pitch + ((x - 4) & ~3);
  const unsigned sh = 8u * (unsigned)((x - 4) & 3);
  int dx0[7], dx1[7], dx2[7], sm0[7], sm1[7], sm2[7];
  int a = 0, b = 0, c = 0;
#pragma unroll
  for (int r = 0; r < 9; ++r) {
    const uint32_t* q = reinterpret_cast<const uint32_t*>(p0 + (size_t)r * pitch);
    const unsigned w0 = q[0], w1 = q[1], w2 = q[2];
    const unsigned lo = __funnelshift_r(w0, w1, sh), hi = __funnelshift_r(w1, w2, sh), top = w2 >> sh;
    int v[9];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      v[j] = (int)((lo >> (8 * j)) & 0xffu);
      v[4 + j] = (int)((hi >> (8 * j)) & 0xffu);
    }
    v[8] = (int)(top & 0xffu);
#pragma unroll
    for (int j = 0; j < 7; ++j) {
      dx0[j] = dx1[j];
      dx1[j] = dx2[j];
      sm0[j] = sm1[j];
      sm1[j] = sm2[j];
      dx2[j] = v[j + 2] - v[j];
      sm2[j] = v[j] + 2 * v[j + 1] + v[j + 2];
    }
    if (r >= 2) {  // rows r-2, r-1, r are live: block row i = r-1
#pragma unroll
      for (int j = 0; j < 7; ++j) {
        int Ix = dx1[j] * 2 + dx0[j] + dx2[j];
        int Iy = sm2[j] - sm0[j];
        a += Ix * Ix;
        b += Iy * Iy;
        c += Ix * Iy;
      }
    }
  }
  float fa = __int2float_rn(a), fb = __int2float_rn(b), fc = __int2float_rn(c);
  float t = __fadd_rn(fa, fb);
  float det = __fsub_rn(__fmul_rn(fa, fb), __fmul_rn(fc, fc));
  float tr = __fmul_rn(__fmul_rn(0.04f, t), t);
  return __fmul_rn(__fsub_rn(det, tr), scale4);
}

// cv::fastAtan2 (scalar float path, degrees), every operation individually rounded.
__device__ __forceinline__ float fast_atan2_deg(float y, float x) {
  const float rad2deg = (float)(180.0 / 3.141592653589793238462643383279502884);
  const float p1 = 0.9997878412794807f * rad2deg;
  const float p3 = -0.3258083974640975f * rad2deg;
  const float p5 = 0.1555786518463281f * rad2deg;
  const float p7 = -0.04432655554792128f * rad2deg;
  const float eps = (float)2.2204460492503131e-16;
  float ax = fabsf(x), ay = fabsf(y);
  float a, c, c2;
  if (ax >= ay) {
    c = __fdiv_rn(ay, __fadd_rn(ax, eps));
    c2 = __fmul_rn(c, c);
    a = __fmul_rn(__fadd_rn(__fmul_rn(__fadd_rn(__fmul_rn(__fadd_rn(__fmul_rn(p7, c2), p5), c2), p3), c2), p1), c);
  } else {
    c = __fdiv_rn(ax, __fadd_rn(ay, eps));
    c2 = __fmul_rn(c, c);
    a = __fsub_rn(
        90.f, __fmul_rn(__fadd_rn(__fmul_rn(__fadd_rn(__fmul_rn(__fadd_rn(__fmul_rn(p7, c2), p5), c2), p3), c2), p1), c));
  }
  if (x < 0.f) a = __fsub_rn(180.f, a);
  if (y < 0.f) a = __fsub_rn(360.f, a);
  return a;
}

// A4 — intensity-centroid angle, one warp per keypoint: lane v+15 sums row v of the 31-px disc.
__device__ float ic_angle_warp(const uint8_t* __restrict__ img, int pitch, int x, int y, int lane) {
  int m10 = 0, m01 = 0;
  if (lane < 31) {
    const int v = lane - 15;
    const int av = v < 0 ? -v : v;
    const int d = (int)((0x3689ABCDDEEEFFFFull >> (4 * av)) & 15ull);  // umax[|v|]
    const uint8_t* row = img + (size_t)(y + v) * pitch + x;
    int s = 0, su = 0;
    for (int u = -d; u <= d; ++u) {
      int val = row[u];
      s += val;
      su += u * val;
    }
    m10 = su;
    m01 = v * s;
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    m10 += __shfl_xor_sync(VSLAM_FULL, m10, o);
    m01 += __shfl_xor_sync(VSLAM_FULL, m01, o);
  }
  return fast_atan2_deg(__int2float_rn(m01), __int2float_rn(m10));
}

__host__ inline size_t select1_smem_bytes(int cap) { return (size_t)cap * (sizeof(uint32_t) + 2 * sizeof(uint16_t)) + 16; }
__host__ inline size_t select2_smem_bytes(int cap) { return (size_t)cap * (sizeof(vrec_t) + 2 * sizeof(uint16_t)) + 16; }

// retainBest(2 * N_l) on the FAST score.  grid = (n_img, nlevels) — level 0 (the longest problems) is scheduled
// first — block = 128: four warps gather the strips, then one warp replays.  Dynamic smem =
// select1_smem_bytes(smem_cap).
// In: per-strip raster-ordered corner lists of k_fast.  Out: sel1[level region][0..m1), sel1_cnt[img][level].
// A level with more corners than smem_cap (or 65535) is replayed in global memory (sel1 as the work array,
// 32-bit index lists in glists) — slower, same result.
__global__ void __launch_bounds__(128) k_select1(const __grid_constant__ PyrP P, int smem_cap,
                                                 const int* __restrict__ strip_cnt,
                                                 const uint32_t* __restrict__ strip_rec, uint32_t* __restrict__ sel1,
                                                 uint32_t* __restrict__ glists, int* __restrict__ sel1_cnt) {
  extern __shared__ __align__(16) uint8_t smem[];
  __shared__ int s_off[129];
  const int img = blockIdx.x, l = blockIdx.y, tid = threadIdx.x, lane = tid & 31;
  const LevelP& L = P.lv[l];
  const int* cnt = strip_cnt + (long long)img * P.total_fast_strips + L.fast_strip0;
  const uint32_t* lrec = strip_rec + (long long)img * P.rec_img_stride + L.rec_off;
  uint32_t* out = sel1 + (long long)img * P.rec_img_stride + L.rec_off;
  // strip offsets (nstrips <= 128)
  if (tid < 32) {
    int run = 0;
    for (int b = 0; b < L.fast_nstrips; b += 32) {
      int i = b + lane;
      int c = i < L.fast_nstrips ? cnt[i] : 0;
      int incl = c;
#pragma unroll
      for (int d = 1; d < 32; d <<= 1) {
        int o = __shfl_up_sync(VSLAM_FULL, incl, d);
        if (lane >= d) incl += o;
      }
      if (i < L.fast_nstrips) s_off[i] = run + incl - c;
      run += __shfl_sync(VSLAM_FULL, incl, 31);
    }
    if (lane == 0) s_off[L.fast_nstrips] = run;
  }
  __syncthreads();
  const int n_in = s_off[L.fast_nstrips];
  const bool use_smem = n_in <= smem_cap && n_in <= 65535;
  uint32_t* rec = use_smem ? reinterpret_cast<uint32_t*>(smem) : out;
  // gather with all 128 threads: one warp per strip at a time, so every load is independent and coalesced
  for (int s0 = tid >> 5; s0 < L.fast_nstrips; s0 += 4) {
    const uint32_t* src = lrec + (long long)s0 * L.strip_cap;
    const int o = s_off[s0], c = s_off[s0 + 1] - o;
    for (int i = lane; i < c; i += 32) rec[o + i] = src[i];
  }
  __syncthreads();
  if (tid >= 32) return;  // the replay itself is one warp; no block barrier below this line
  int m;
  if (use_smem) {
    uint16_t* l16 = reinterpret_cast<uint16_t*>(smem + (size_t)smem_cap * sizeof(uint32_t));
    m = warp_retain_best<uint32_t, uint16_t>(rec, n_in, 2 * L.nfeat, l16, l16 + smem_cap, lane);
    for (int i = lane; i < m; i += 32) out[i] = rec[i];
  } else {
    uint32_t* l32 = glists + 2 * ((long long)img * P.rec_img_stride + L.rec_off);
    m = warp_retain_best<uint32_t, uint32_t>(rec, n_in, 2 * L.nfeat, l32, l32 + (long long)L.fast_nstrips * L.strip_cap,
                                             lane);
  }
  if (lane == 0) sel1_cnt[(long long)img * P.nlevels + l] = m;
}

// Harris response of every select1 survivor: grid = (blocks, nlevels, n_img), block = 128, thread per survivor.
#ifndef VSLAM_HARRIS_MINB
#define VSLAM_HARRIS_MINB 1
#endif
__global__ void __launch_bounds__(128, VSLAM_HARRIS_MINB) k_harris(const __grid_constant__ PyrP P, const uint8_t* __restrict__ in_l,
                                                const uint8_t* __restrict__ in_r, const uint8_t* __restrict__ pyr,
                                                const uint32_t* __restrict__ sel1, const int* __restrict__ sel1_cnt,
                                                vrec_t* __restrict__ cand, float harris_scale4) {
  const int l = blockIdx.y, img = blockIdx.z;
  const int m = sel1_cnt[(long long)img * P.nlevels + l];
  const LevelP& L = P.lv[l];
  int pitch;
  const uint8_t* img_base = level_ptr(P, l, img, in_l, in_r, pyr, pitch);
  const long long base = (long long)img * P.rec_img_stride + L.rec_off;
  for (int i = blockIdx.x * 128 + threadIdx.x; i < m; i += gridDim.x * 128) {
    const uint32_t r = sel1[base + i];
    const int x = (int)(r & 0xfffu), y = (int)((r >> 12) & 0xfffu);
    const float h = harris_response(img_base, pitch, x, y, harris_scale4);
    cand[base + i] = ((vrec_t)__float_as_uint(h) << 32) | ((unsigned)y << 16) | (unsigned)x;
  }
}

// retainBest(N_l) on the Harris response.  grid = (n_img, nlevels), block = 32.  Dynamic smem =
// select2_smem_bytes(smem_cap).  Out: lvl_kp[img][level][0..m2), lvl_cnt[img][level].
__global__ void __launch_bounds__(32) k_select2(const __grid_constant__ PyrP P, int smem_cap,
                                                vrec_t* __restrict__ cand, const int* __restrict__ sel1_cnt,
                                                uint32_t* __restrict__ glists, LvlKp* __restrict__ lvl_kp,
                                                int* __restrict__ lvl_cnt, const FrameLayout F) {
  extern __shared__ __align__(16) uint8_t smem[];
  const int img = blockIdx.x, l = blockIdx.y, lane = threadIdx.x;
  const LevelP& L = P.lv[l];
  const int m1 = sel1_cnt[(long long)img * P.nlevels + l];
  vrec_t* gl = cand + (long long)img * P.rec_img_stride + L.rec_off;
  const bool use_smem = m1 <= smem_cap && m1 <= 65535;
  vrec_t* rec = use_smem ? reinterpret_cast<vrec_t*>(smem) : gl;
  if (use_smem)
    for (int i = lane; i < m1; i += 32) rec[i] = gl[i];
  __syncwarp();
  int m2;
  if (use_smem) {
    uint16_t* l16 = reinterpret_cast<uint16_t*>(smem + (size_t)smem_cap * sizeof(vrec_t));
    m2 = warp_retain_best<vrec_t, uint16_t>(rec, m1, L.nfeat, l16, l16 + smem_cap, lane);
  } else {
    uint32_t* l32 = glists + 2 * ((long long)img * P.rec_img_stride + L.rec_off);
    m2 = warp_retain_best<vrec_t, uint32_t>(rec, m1, L.nfeat, l32, l32 + (long long)L.fast_nstrips * L.strip_cap, lane);
  }
  if (m2 > P.kp_cap) {  // more ties than the handle was sized for: flag, keep the first kp_cap
    if (lane == 0) atomicMax(&F.hdr[img / P.eyes].status, 4 /* VSLAM_ERR_KEYPOINT_OVERFLOW */);
    m2 = P.kp_cap;
  }
  LvlKp* out = lvl_kp + ((long long)img * P.nlevels + l) * P.kp_cap;
  for (int i = lane; i < m2; i += 32) {
    const vrec_t r = rec[i];
    LvlKp k;
    k.xy = (unsigned)r;
    k.resp = __uint_as_float((unsigned)(r >> 32));
    out[i] = k;
  }
  if (lane == 0) lvl_cnt[(long long)img * P.nlevels + l] = m2;
}

// A3, fused: retainBest(2 N_l) on the FAST score -> HarrisResponses of the survivors -> retainBest(N_l) on the Harris
// response, one block of 128 threads per (image, level), every phase spread over the four warps.  The three-kernel
// form above costs two grid-wide dependencies per chunk (k_harris waits for the slowest k_select1 block, k_select2
// for the slowest k_harris block); here a level's chain runs back to back.
// grid = (n_img, nlevels), block = 128.  Dynamic smem = select_fused_smem_bytes(cap1, cap2): cap1 FAST-score records
// (+ index lists) for the first replay, cap2 Harris records for the second; a problem larger than its budget is
// replayed in global memory (sel1 / cand as the work arrays, glists for the index lists) — slower, same result.
__host__ inline size_t select_fused_smem_bytes(int cap1, int cap2) {
  const size_t a = (size_t)cap1 * (sizeof(uint32_t) + 2 * sizeof(uint16_t));
  const size_t b = (size_t)cap2 * (sizeof(vrec_t) + 2 * sizeof(uint16_t));
  return (a > b ? a : b) + 16;
}
__global__ void __launch_bounds__(128) k_select_fused(const __grid_constant__ PyrP P, int cap1, int cap2,
                                                      const uint8_t* __restrict__ in_l, const uint8_t* __restrict__ in_r,
                                                      const uint8_t* __restrict__ pyr, const int* __restrict__ strip_cnt,
                                                      const uint32_t* __restrict__ strip_rec, uint32_t* __restrict__ sel1,
                                                      vrec_t* __restrict__ cand, uint32_t* __restrict__ glists,
                                                      int* __restrict__ sel1_cnt, LvlKp* __restrict__ lvl_kp,
                                                      int* __restrict__ lvl_cnt, float harris_scale4, const FrameLayout F) {
  extern __shared__ __align__(16) uint8_t smem[];
  __shared__ int s_off[129];
  __shared__ int s_w[12];
  const int img = blockIdx.x, l = blockIdx.y, tid = threadIdx.x, lane = tid & 31;
  const LevelP& L = P.lv[l];
  const int* cnt = strip_cnt + (long long)img * P.total_fast_strips + L.fast_strip0;
  const long long base = (long long)img * P.rec_img_stride + L.rec_off;
  const uint32_t* lrec = strip_rec + base;
  uint32_t* out1 = sel1 + base;
  uint32_t* gl32 = glists + 2 * base;
  const long long gl_half = (long long)L.fast_nstrips * L.strip_cap;
  if (tid < 32) {  // strip offsets (nstrips <= 128)
    int run = 0;
    for (int b = 0; b < L.fast_nstrips; b += 32) {
      const int i = b + lane;
      const int c = i < L.fast_nstrips ? cnt[i] : 0;
      int incl = c;
#pragma unroll
      for (int d = 1; d < 32; d <<= 1) {
        const int o = __shfl_up_sync(VSLAM_FULL, incl, d);
        if (lane >= d) incl += o;
      }
      if (i < L.fast_nstrips) s_off[i] = run + incl - c;
      run += __shfl_sync(VSLAM_FULL, incl, 31);
    }
    if (lane == 0) s_off[L.fast_nstrips] = run;
  }
  __syncthreads();
  const int n_in = s_off[L.fast_nstrips];
  // ---- gather: element e of the level's raster-ordered list lives in strip s with s_off[s] <= e < s_off[s+1]; four
  //      independent loads per thread are in flight
  const bool smem1 = n_in <= cap1 && n_in <= 65535;
  uint32_t* rec = smem1 ? reinterpret_cast<uint32_t*>(smem) : out1;
  {
    int st = 0;
    for (int e0 = tid; e0 < n_in; e0 += 4 * 128) {
      uint32_t v[4];
      int st_k = st;
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        const int e = e0 + k * 128;
        if (e < n_in) {
          while (s_off[st_k + 1] <= e) ++st_k;
          v[k] = lrec[(long long)st_k * L.strip_cap + (e - s_off[st_k])];
        }
      }
      st = st_k;
#pragma unroll
      for (int k = 0; k < 4; ++k)
        if (e0 + k * 128 < n_in) rec[e0 + k * 128] = v[k];
    }
  }
  __syncthreads();
  // ---- retainBest(2 N_l) on the FAST score
  int m1;
  if (smem1) {
    uint16_t* l16 = reinterpret_cast<uint16_t*>(smem + (size_t)cap1 * sizeof(uint32_t));
    m1 = block_retain_best<uint32_t, uint16_t>(rec, n_in, 2 * L.nfeat, l16, l16 + cap1, s_w);
    for (int i = tid; i < m1; i += 128) out1[i] = rec[i];
  } else {
    m1 = block_retain_best<uint32_t, uint32_t>(rec, n_in, 2 * L.nfeat, gl32, gl32 + gl_half, s_w);
  }
  if (tid == 0) sel1_cnt[(long long)img * P.nlevels + l] = m1;
  __syncthreads();  // out1[0 .. m1) is visible to the whole block; the first replay's shared memory is free
  // ---- HarrisResponses of the survivors (thread per survivor)
  int pitch;
  const uint8_t* img_base = level_ptr(P, l, img, in_l, in_r, pyr, pitch);
  const bool smem2 = m1 <= cap2 && m1 <= 65535;
  vrec_t* gcand = cand + base;
  vrec_t* rec2 = smem2 ? reinterpret_cast<vrec_t*>(smem) : gcand;
  for (int i = tid; i < m1; i += 128) {
    const uint32_t r = out1[i];
    const int x = (int)(r & 0xfffu), y = (int)((r >> 12) & 0xfffu);
    const float hr = harris_response(img_base, pitch, x, y, harris_scale4);
    rec2[i] = ((vrec_t)__float_as_uint(hr) << 32) | ((unsigned)y << 16) | (unsigned)x;
  }
  __syncthreads();
  // ---- retainBest(N_l) on the Harris response
  int m2;
  if (smem2) {
    uint16_t* l16 = reinterpret_cast<uint16_t*>(smem + (size_t)cap2 * sizeof(vrec_t));
    m2 = block_retain_best<vrec_t, uint16_t>(rec2, m1, L.nfeat, l16, l16 + cap2, s_w);
  } else {
    m2 = block_retain_best<vrec_t, uint32_t>(rec2, m1, L.nfeat, gl32, gl32 + gl_half, s_w);
  }
  if (m2 > P.kp_cap) {  // more ties than the handle was sized for: flag, keep the first kp_cap
    if (tid == 0) atomicMax(&F.hdr[img / P.eyes].status, 4 /* VSLAM_ERR_KEYPOINT_OVERFLOW */);
    m2 = P.kp_cap;
  }
  LvlKp* out = lvl_kp + ((long long)img * P.nlevels + l) * P.kp_cap;
  for (int i = tid; i < m2; i += 128) {
    const vrec_t r = rec2[i];
    LvlKp k;
    k.xy = (unsigned)r;
    k.resp = __uint_as_float((unsigned)(r >> 32));
    out[i] = k;
  }
  if (tid == 0) lvl_cnt[(long long)img * P.nlevels + l] = m2;
}

// ----------------------------------------------------------------------------------------------------------
// A5 — the blur cv::ORB applies to each level: separable 7-tap float32 Gaussian (sigma 2) as OpenCV's AVX2
// sepFilter2D evaluates it: row pass s = k0*x0, s = fma(x_j, k_j, s); column pass s = k3*r3,
// s = fma(r_{3+j} + r_{3-j}, k_{3+j}, s); BORDER_REFLECT_101; round-half-even to u8.
//
// Register-resident sliding window, no shared memory, no barriers.  One LANE owns an 8-column x (<= 64)-row item of
// one level and walks down its rows: per row it loads the 16 source bytes x-4 .. x+11 (four aligned words), converts
// them to f32, evaluates the row pass for its 8 columns and pushes the result into a 7-row window held in registers;
// the column pass of output row y-3 reads that window.  All float work is PACKED f32x2 (FFMA2 / FMUL2 / FADD2, two
// IEEE-rounded operations per instruction): a pair is (column k, column k+4), k = 0..3, so the row-pass operands
// (f[a], f[a+4]) and the column-pass operands stay aligned in the same 64-bit registers from conversion to store.
// grid = (ceil(total_blur_items / 128), n_img), block = 128; a level's items are padded to a multiple of 32, so a
// warp never spans two levels.  Items are numbered column-group fastest: adjacent lanes read adjacent 8-byte chunks.
// ----------------------------------------------------------------------------------------------------------
__device__ __forceinline__ int reflect101(int i, int n) {
  if (n == 1) return 0;
  while (i < 0 || i >= n) i = i < 0 ? -i : 2 * n - 2 - i;
  return i;
}

typedef unsigned long long f32x2_t;
__device__ __forceinline__ f32x2_t pk2(float lo, float hi) {
  f32x2_t r;
  asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi));
  return r;
}
__device__ __forceinline__ f32x2_t pk2u(unsigned lo, unsigned hi) {
  f32x2_t r;
  asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "r"(lo), "r"(hi));
  return r;
}
__device__ __forceinline__ void upk2u(f32x2_t v, unsigned& lo, unsigned& hi) {
  asm("mov.b64 {%0, %1}, %2;" : "=r"(lo), "=r"(hi) : "l"(v));
}
__device__ __forceinline__ void upk2f(f32x2_t v, float& lo, float& hi) {
  asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v));
}
__device__ __forceinline__ f32x2_t fma2(f32x2_t a, f32x2_t b, f32x2_t c) {
  f32x2_t d;
  asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c));
  return d;
}
__device__ __forceinline__ f32x2_t mul2(f32x2_t a, f32x2_t b) {
  f32x2_t d;
  asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b));
  return d;
}
__device__ __forceinline__ f32x2_t add2(f32x2_t a, f32x2_t b) {
  f32x2_t d;
  asm("add.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b));
  return d;
}

#ifndef VSLAM_BLUR_RING
#define VSLAM_BLUR_RING 8
#endif
constexpr int kBlurRing = VSLAM_BLUR_RING;  // source rows in flight per lane (cp.async ring in shared memory), power of 2
__device__ __forceinline__ void blur_item(const PyrP& P, const int img, const int item, const uint8_t* __restrict__ in_l,
                                          const uint8_t* __restrict__ in_r, const uint8_t* __restrict__ pyr,
                                          uint8_t* __restrict__ blur) {
  int l = 0;
  for (int k = 1; k < P.nlevels; ++k)
    if (item >= P.lv[k].blur_item0) l = k;  // warp-uniform
  const LevelP& L = P.lv[l];
  const int li = item - L.blur_item0;
  const int strip = li / L.blur_ncg, cg = li - strip * L.blur_ncg;
  if (strip >= L.blur_nstr) return;  // padding lanes (no warp collectives below)
  const int x = 8 * (cg + L.blur_cg0), y0 = L.blur_y0 + strip * L.blur_rs;
  const int rs = min(L.blur_rs, L.blur_y1 - y0);  // output rows of this item
  const int nblk = (L.blur_rs + 6) / 7;     // main-loop trips of 7 output rows (warp-uniform)
  int spitch;
  const uint8_t* src = level_ptr(P, l, img, in_l, in_r, pyr, spitch);
  uint8_t* dst = blur + (long long)img * P.blur_img_stride + L.blur_off + (size_t)y0 * L.pitch + x;

  // ---- per-lane window geometry.  Word k nominally holds pixels x-4+4k .. x-1+4k.  Bytes left of pixel 0 are
  //      reflected out of word 0 (pixels 0..3: the address clamp loads exactly that word); bytes right of the last
  //      pixel e are reflected out of the strip S = pixels e-3 .. e.  One PRMT per word applies both (identity
  //      selector for interior lanes: the integer pipe has slack, a divergent branch would cost more).
  const int e = L.w - 1, xlast = e & ~3;
  const uint8_t* pw[4];
  unsigned sel[4];
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    const int xk = x - 4 + 4 * k;
    pw[k] = src + min(max(xk, 0), xlast);
    unsigned sk = 0;
#pragma unroll
    for (int b = 0; b < 4; ++b) {
      const int px = xk + b, d = px - e;
      const unsigned nib = px < 0 ? (unsigned)(-px) & 3u : (d <= 0 ? (unsigned)b : (d <= 3 ? (unsigned)(7 - d) : 0u));
      sk |= nib << (4 * b);
    }
    sel[k] = sk;
  }
  const bool redge = x + 11 > e;
  const unsigned sh = 8u * (unsigned)((e - 3) & 3);
  const uint8_t* ps = src + ((e - 3) & ~3);
  const int ps2 = sh ? 4 : 0;
  int hm1 = L.h - 1, dpitch = L.pitch;
  // Loop invariants must live in registers: under register pressure ptxas otherwise re-derives them inside the row
  // loop through indexed constant-bank loads (LDC c[0x0][R+..], long-scoreboard latency in front of every SHF / PRMT).
  unsigned shv = sh;
  asm volatile("" : "+r"(hm1), "+r"(dpitch), "+r"(shv), "+r"(spitch));
  asm volatile("" : "+r"(sel[0]), "+r"(sel[1]), "+r"(sel[2]), "+r"(sel[3]));
  asm volatile("" : "+l"(dst));

  // Source rows reach the lane through a private ring in shared memory filled by cp.async (LDGSTS: global -> shared
  // without a register, kBlurRing rows in flight per lane — enough for a DRAM miss, which two rows of register
  // prefetch were not).  Per row: 4 window words (16 B, read back as one conflict-free LDS.128) + for right-edge lanes
  // the 2 words that hold S.  issue_row(t) enqueues source row t into slot t % kBlurRing as one commit group;
  // take_row() waits for the oldest group and applies the border selectors.
  __shared__ __align__(16) uint4 s_ring[kBlurRing][128];
  __shared__ __align__(8) uint2 s_ringS[kBlurRing][128];
  const unsigned ring0 = (unsigned)__cvta_generic_to_shared(&s_ring[0][threadIdx.x]);
  const unsigned ringS0 = (unsigned)__cvta_generic_to_shared(&s_ringS[0][threadIdx.x]);
  auto issue_row = [&](int t, int slot) {
    // reflect-101 of yy = y0 - 3 + t (|yy| mirrors the top, (h-1) - |(h-1) - .| the bottom), clamped for the rows
    // past the mirror that feed no stored output
    const int gy = max(hm1 - abs(hm1 - abs(y0 - 3 + t)), 0);
    const unsigned ro = (unsigned)gy * (unsigned)spitch;
    const unsigned d = ring0 + (unsigned)slot * (unsigned)(128 * sizeof(uint4));
#pragma unroll
    for (int k = 0; k < 4; ++k)
      asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(d + 4u * k), "l"(pw[k] + ro) : "memory");
    if (redge) {
      const unsigned ds = ringS0 + (unsigned)slot * (unsigned)(128 * sizeof(uint2));
      asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(ds), "l"(ps + ro) : "memory");
      asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(ds + 4u), "l"(ps + ro + ps2) : "memory");
    }
    asm volatile("cp.async.commit_group;" ::: "memory");
  };
  auto take_row = [&](int slot, unsigned (&w)[4]) {
    asm volatile("cp.async.wait_group %0;" ::"n"(kBlurRing - 1) : "memory");
    uint4 r;
    asm volatile("ld.shared.v4.u32 {%0, %1, %2, %3}, [%4];"
                 : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w)
                 : "r"(ring0 + (unsigned)slot * (unsigned)(128 * sizeof(uint4))));
    unsigned sA = 0, sB = 0;
    if (redge)
      asm volatile("ld.shared.v2.u32 {%0, %1}, [%2];"
                   : "=r"(sA), "=r"(sB)
                   : "r"(ringS0 + (unsigned)slot * (unsigned)(128 * sizeof(uint2))));
    const unsigned S = __funnelshift_r(sA, sB, shv);
    w[0] = __byte_perm(r.x, S, sel[0]);
    w[1] = __byte_perm(r.y, S, sel[1]);
    w[2] = __byte_perm(r.z, S, sel[2]);
    w[3] = __byte_perm(r.w, S, sel[3]);
  };

  const f32x2_t K0 = pk2(0x1.1f5f62p-4f, 0x1.1f5f62p-4f), K1 = pk2(0x1.0c70fcp-3f, 0x1.0c70fcp-3f);
  const f32x2_t K2 = pk2(0x1.869472p-3f, 0x1.869472p-3f), K3 = pk2(0x1.ba95c0p-3f, 0x1.ba95c0p-3f);
  const f32x2_t NEG23 = pk2(-8388608.0f, -8388608.0f), MAGIC = pk2(12582912.0f, 12582912.0f);

  // Row pass of one source row (window words w) into r[0..3].
  // u8 -> f32: the byte goes into the mantissa of 2^23, then (2^23 + b) - 2^23 is exact.  f[a] = pixel x-3+a is
  // byte a+1 of the window; pair a = (f[a], f[a+4]).
  auto row_pass = [&](const unsigned (&w)[4], f32x2_t (&r)[4]) {
    f32x2_t p[10];
#pragma unroll
    for (int a = 0; a < 10; ++a) {
      const unsigned s4 = 0x7540u | (unsigned)((a + 1) & 3);
      const unsigned lo = __byte_perm(w[(a + 1) >> 2], 0x4B000000u, s4);
      const unsigned hi = __byte_perm(w[(a + 5) >> 2], 0x4B000000u, s4);
      p[a] = add2(pk2u(lo, hi), NEG23);
    }
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      f32x2_t s = mul2(p[k], K0);
      s = fma2(p[k + 1], K1, s);
      s = fma2(p[k + 2], K2, s);
      s = fma2(p[k + 3], K3, s);
      s = fma2(p[k + 4], K2, s);
      s = fma2(p[k + 5], K1, s);
      s = fma2(p[k + 6], K0, s);
      r[k] = s;
    }
  };

  f32x2_t win[7][4];  // row-pass results of the last 7 source rows; every index is a compile-time constant
  static_assert((kBlurRing & (kBlurRing - 1)) == 0 && kBlurRing >= 2, "ring size must be a power of two");
#pragma unroll
  for (int t = 0; t < kBlurRing; ++t) issue_row(t, t);
#pragma unroll
  for (int t = 0; t < 6; ++t) {  // fill the window: source rows 0..5 of the item
    unsigned cw[4];
    take_row(t & (kBlurRing - 1), cw);
    row_pass(cw, win[t]);
    issue_row(t + kBlurRing, t & (kBlurRing - 1));  // the slot is free once its words sit in registers
  }
  int slot = 6 & (kBlurRing - 1);  // ring slot of source row t, carried because 7 rows per trip do not divide the ring
#pragma unroll 1
  for (int blk = 0; blk < nblk; ++blk) {
    const int tb = 6 + 7 * blk;
#pragma unroll
    for (int ph = 0; ph < 7; ++ph) {
      const int t = tb + ph;          // source row of the item; window slot (6 + ph) % 7
      unsigned cw[4];
      take_row(slot, cw);
      row_pass(cw, win[(6 + ph) % 7]);
      issue_row(t + kBlurRing, slot);
      slot = (slot + 1) & (kBlurRing - 1);
      // column pass of output row t - 6: source rows t-6 .. t live in slots (ph + j) % 7, j = 0..6
      unsigned b[8];
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        f32x2_t s = mul2(win[(ph + 3) % 7][k], K3);
        s = fma2(add2(win[(ph + 4) % 7][k], win[(ph + 2) % 7][k]), K2, s);
        s = fma2(add2(win[(ph + 5) % 7][k], win[(ph + 1) % 7][k]), K1, s);
        s = fma2(add2(win[(ph + 6) % 7][k], win[(ph + 0) % 7][k]), K0, s);
        upk2u(add2(s, MAGIC), b[k], b[k + 4]);  // low byte = round-half-even(s), 0 <= s < 256
      }
      uint2 o;
      o.x = __byte_perm(__byte_perm(b[0], b[1], 0x0040), __byte_perm(b[2], b[3], 0x0040), 0x5410);
      o.y = __byte_perm(__byte_perm(b[4], b[5], 0x0040), __byte_perm(b[6], b[7], 0x0040), 0x5410);
      if (t - 6 < rs) *reinterpret_cast<uint2*>(dst + (unsigned)(t - 6) * (unsigned)dpitch) = o;
    }
  }
  asm volatile("cp.async.wait_group 0;" ::: "memory");  // drain the look-ahead rows before the block's memory is reused
}

// Persistent launch: grid = (blocks, 1) with a fixed number of blocks per SM chosen by the host; every block walks
// the (image, 128-item group) space with a grid stride.  A blur that cannot fill the SMs on its own leaves registers
// and shared memory for the FAST blocks of the same chunk: the two kernels are bound by different pipes (FMA vs ALU)
// and are launched together on two streams.
__global__ void __launch_bounds__(128, VSLAM_BLUR_MINB) k_blur(const __grid_constant__ PyrP P, int nimg,
                                                               const uint8_t* __restrict__ in_l,
                                                               const uint8_t* __restrict__ in_r,
                                                               const uint8_t* __restrict__ pyr,
                                                               uint8_t* __restrict__ blur) {
  const int groups = (P.total_blur_items + 127) >> 7;  // 128-item groups per image
  const int total = nimg * groups;
  for (int gi = blockIdx.x; gi < total; gi += gridDim.x) {
    const int img = gi / groups, grp = gi - img * groups;
    blur_item(P, img, grp * 128 + (int)threadIdx.x, in_l, in_r, pyr, blur);
  }
}

// ----------------------------------------------------------------------------------------------------------
// A6 — steered BRIEF + KeyPoint assembly.  grid = (ceil(kp_cap / 4), n_img), block = 128 (4 warps, one keypoint
// each); lane b computes descriptor byte b (16 rotated pattern points).  Output order: levels 0..n-1 concatenated.
// ----------------------------------------------------------------------------------------------------------
// (float)cos((double)x), (float)sin((double)x) for 0 <= x < 2*pi + eps, as OpenCV's steering computes them with the
// host libm.  Double-precision evaluation (Cody-Waite reduction by pi/2, Cephes minimax kernels, <= ~1 ulp(double))
// so that the value rounded to float equals the correctly rounded one except with probability ~1e-8 per call.
// Written out instead of calling cos()/sin(): the CUDA library versions keep their coefficient table in local
// memory, which cost 112 B of stack stores per thread at kernel entry.
__device__ __forceinline__ void sincos_f64_to_f32(float xf, float& cf, float& sf) {
  const double x = (double)xf;
  const double q = rint(x * 6.36619772367581382433e-01);
  double t = fma(-q, 1.57079632679489655800e+00, x);
  t = fma(-q, 6.12323399573676603587e-17, t);
  const double z = t * t;
  double ps = 1.58962301576546568060e-10;
  ps = fma(ps, z, -2.50507477628578072866e-8);
  ps = fma(ps, z, 2.75573136213857245213e-6);
  ps = fma(ps, z, -1.98412698295895385996e-4);
  ps = fma(ps, z, 8.33333333332211858878e-3);
  ps = fma(ps, z, -1.66666666666666307295e-1);
  const double sn = fma(t * z, ps, t);
  double pc = -1.13596475577881948265e-11;
  pc = fma(pc, z, 2.08757232129817482790e-9);
  pc = fma(pc, z, -2.75573143513906633035e-7);
  pc = fma(pc, z, 2.48015872894767294178e-5);
  pc = fma(pc, z, -1.38888888888741095749e-3);
  pc = fma(pc, z, 4.16666666666666019037e-2);
  const double cs = fma(z * z, pc, fma(z, -0.5, 1.0));
  const int iq = (int)q;
  const double s1 = (iq & 1) ? cs : sn, c1 = (iq & 1) ? sn : cs;
  sf = (float)((iq & 2) ? -s1 : s1);
  cf = (float)(((iq + 1) & 2) ? -c1 : c1);
}

struct __align__(4) KeyPoint28 {
  float x, y, size, angle, response;
  int octave, class_id;
};

constexpr int kBriefR = 18;                     // max |rotated pattern coordinate| (13*sqrt(2) = 18.38 -> 18)
constexpr int kBriefRows = 2 * kBriefR + 1;     // 37
constexpr int kBriefRowWords = 11;              // smem row stride in words (10 loaded + 1 pad; 11 is coprime with 32)
constexpr int kAngleR = 15;
constexpr int kAngleRows = 2 * kAngleR + 1;     // 31
constexpr int kAngleRowWords = 9;               // 31 px + up to 3 alignment bytes = 9 words
constexpr int kPatchWords = kBriefRows * kBriefRowWords + kAngleRows * kAngleRowWords;

#ifndef VSLAM_BRIEF_PACKED
#define VSLAM_BRIEF_PACKED 1  // steering arithmetic on packed f32x2 (both points of a test per instruction)
#endif
#ifndef VSLAM_BRIEF_WARPS
#define VSLAM_BRIEF_WARPS 4
#endif
constexpr int kBriefWarps = VSLAM_BRIEF_WARPS;  // keypoints (warps) per block
// Descriptor byte `lane` of a keypoint: tests 8*lane .. 8*lane+7 of the rotated pattern.  `center` is the shared-window
// address of the patch centre with the cvRound bias folded in (see k_brief), ROWB the patch row stride in bytes.
// a = (float)cos, b = (float)sin.  ix = cvRound(x*a - y*b), iy = cvRound(x*b + y*a), every product and sum rounded to f32.
template <unsigned ROWB>
__device__ __forceinline__ unsigned steered_brief_byte(const unsigned center, const float a, const float b, const int lane) {
  unsigned byte = 0;
#if VSLAM_BRIEF_PACKED
  // Packed f32x2 where it is exact: the pattern bytes (biased by 128) are dropped into the mantissa of 2^23 by PRMT and
  // un-biased by one packed add per two coordinates; the four products of a test's two points take four FMUL2
  // (x*a - y*b == x*a + y*(-b)); the cvRound bias of two coordinates is one FADD2.  The sums of two products stay
  // SCALAR: ptxas (12.9) contracts mul.rn.f32x2 + add.rn.f32x2 into one FFMA2 - a single rounding, not what the
  // reference computes - and honours .rn only on the scalar forms.
  const f32x2_t AA = pk2(a, a), BB = pk2(b, b), NB = pk2(-b, -b), MAGIC = pk2(12582912.0f, 12582912.0f),
                 UNBIAS = pk2(-8388736.0f, -8388736.0f);
#pragma unroll
  for (int q = 0; q < 8; ++q) {
    const unsigned pw = g_brief_pattern_i8[q][lane];  // test 8*lane + q: (x0, y0, x1, y1), bytes biased by 128
    const f32x2_t X = add2(pk2u(__byte_perm(pw, 0x4B000000u, 0x7440), __byte_perm(pw, 0x4B000000u, 0x7442)), UNBIAS);
    const f32x2_t Y = add2(pk2u(__byte_perm(pw, 0x4B000000u, 0x7441), __byte_perm(pw, 0x4B000000u, 0x7443)), UNBIAS);
    unsigned ix0, ix1, iy0, iy1;
    float xa0, xa1, ynb0, ynb1, xb0, xb1, ya0, ya1;
    upk2f(mul2(X, AA), xa0, xa1);
    upk2f(mul2(Y, NB), ynb0, ynb1);
    upk2f(mul2(X, BB), xb0, xb1);
    upk2f(mul2(Y, AA), ya0, ya1);
    upk2u(add2(pk2(__fadd_rn(xa0, ynb0), __fadd_rn(xa1, ynb1)), MAGIC), ix0, ix1);
    upk2u(add2(pk2(__fadd_rn(xb0, ya0), __fadd_rn(xb1, ya1)), MAGIC), iy0, iy1);
    const unsigned t0 = lds8(center + iy0 * ROWB + ix0), t1 = lds8(center + iy1 * ROWB + ix1);
    byte |= (unsigned)(t0 < t1) << q;
  }
#else
  auto raw = [](float v) -> unsigned { return __float_as_uint(__fadd_rn(v, 12582912.0f)); };
#pragma unroll
  for (int q = 0; q < 8; ++q) {
    const unsigned pw = g_brief_pattern_i8[q][lane];  // test 8*lane + q: (x0, y0, x1, y1) as int8, exact in f32
    float4 t;
    t.x = (float)(int)(signed char)(pw & 0xffu);
    t.y = (float)(int)(signed char)((pw >> 8) & 0xffu);
    t.z = (float)(int)(signed char)((pw >> 16) & 0xffu);
    t.w = (float)(int)(signed char)(pw >> 24);
    const unsigned ix0 = raw(__fsub_rn(__fmul_rn(t.x, a), __fmul_rn(t.y, b)));
    const unsigned iy0 = raw(__fadd_rn(__fmul_rn(t.x, b), __fmul_rn(t.y, a)));
    const unsigned ix1 = raw(__fsub_rn(__fmul_rn(t.z, a), __fmul_rn(t.w, b)));
    const unsigned iy1 = raw(__fadd_rn(__fmul_rn(t.z, b), __fmul_rn(t.w, a)));
    const unsigned t0 = lds8(center + iy0 * ROWB + ix0), t1 = lds8(center + iy1 * ROWB + ix1);
    byte |= (unsigned)(t0 < t1) << q;
  }
#endif
  return byte;
}

__global__ void __launch_bounds__(32 * kBriefWarps) k_brief(const __grid_constant__ PyrP P, const uint8_t* __restrict__ in_l,
                                               const uint8_t* __restrict__ in_r, const uint8_t* __restrict__ pyr,
                                               const uint8_t* __restrict__ blur, const LvlKp* __restrict__ lvl_kp,
                                               const int* __restrict__ lvl_cnt, const FrameLayout F) {
  __shared__ uint32_t s_patch[kBriefWarps][kPatchWords];
  const int img = blockIdx.y, lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int j = blockIdx.x * kBriefWarps + warp;
  const int frame = P.eyes == 2 ? img >> 1 : img, eye = P.eyes == 2 ? (img & 1) : 0;  // eyes is 1 or 2
  const int* cnt = lvl_cnt + (long long)img * P.nlevels;
  int total = 0, l = -1, jl = 0;
  {
    int c[VSLAM_MAX_LEVELS];  // all loads first (independent), then the prefix walk
#pragma unroll
    for (int k = 0; k < VSLAM_MAX_LEVELS; ++k) c[k] = k < P.nlevels ? cnt[k] : 0;
#pragma unroll
    for (int k = 0; k < VSLAM_MAX_LEVELS; ++k) {
      if (l < 0 && j < total + c[k]) {
        l = k;
        jl = j - total;
      }
      total += c[k];
    }
  }
  ResHdr* hdr = F.hdr + frame;
  if (blockIdx.x == 0 && threadIdx.x == 0) {
    hdr->n_kp[eye] = min(total, P.kp_cap);
    if (total > P.kp_cap) atomicMax(&hdr->status, 4 /* VSLAM_ERR_KEYPOINT_OVERFLOW */);
  }
  if (l < 0 || j >= P.kp_cap) return;  // warp-uniform
  const LevelP& L = P.lv[l];
  const LvlKp kp = lvl_kp[((long long)img * P.nlevels + l) * P.kp_cap + jl];
  const int cx = (int)(kp.xy & 0xffffu), cy = (int)(kp.xy >> 16);

  // ---- stage the two patches this keypoint reads with aligned 32-bit loads (2 sectors per row instead of one
  //      sector per sample): 37x37 of the blurred level (BRIEF) and 31x31 of the un-blurred level (orientation).
  //      Keypoints are >= 31 px from every edge (runByImageBorder), so both windows are inside the image.
  uint32_t* sb = s_patch[warp];
  uint32_t* sa = sb + kBriefRows * kBriefRowWords;
  int upitch;
  const uint8_t* unblurred = level_ptr(P, l, img, in_l, in_r, pyr, upitch);
  const uint8_t* blurred = blur + (long long)img * P.blur_img_stride + L.blur_off;
  const int xb0 = (cx - kBriefR) & ~3, xa0 = (cx - kAngleR) & ~3;
  {
    // three rows per step: lanes 0-9 / 10-19 / 20-29 copy the 10 words of a BRIEF row, lanes 0-8 / 9-17 / 18-26 the
    // 9 words of an orientation row, with cp.async (global -> shared without a register round trip; all 24 copies of
    // a lane are in flight at once).  Only the last step of each patch has rows past the end (row 37 / 38, 31 / 32).
    const int rb = lane / 10, wb = lane - 10 * rb;
    const int ra = lane / 9, wa = lane - 9 * ra;
    const uint8_t* pb = blurred + (size_t)(cy - kBriefR + rb) * L.pitch + xb0 + 4 * wb;
    const uint8_t* pa = unblurred + (size_t)(cy - kAngleR + ra) * upitch + xa0 + 4 * wa;
    constexpr int NB = (kBriefRows + 2) / 3, NA = (kAngleRows + 2) / 3;  // 13, 11 steps of three rows
    static_assert(3 * (NB - 1) == kBriefRows - 1 && 3 * (NA - 1) == kAngleRows - 1, "last step holds exactly one row");
    const size_t stepb = (size_t)3 * L.pitch, stepa = (size_t)3 * upitch;
    const unsigned qb = (unsigned)__cvta_generic_to_shared(sb + rb * kBriefRowWords + wb);
    const unsigned qa = (unsigned)__cvta_generic_to_shared(sa + ra * kAngleRowWords + wa);
    if (lane < 30) {
#pragma unroll
      for (int k = 0; k < NB; ++k) {
        if (k < NB - 1 || rb == 0)
          asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(qb + (unsigned)(k * 3 * kBriefRowWords * 4)), "l"(pb)
                       : "memory");
        pb += stepb;
      }
    }
    if (lane < 27) {
#pragma unroll
      for (int k = 0; k < NA; ++k) {
        if (k < NA - 1 || ra == 0)
          asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(qa + (unsigned)(k * 3 * kAngleRowWords * 4)), "l"(pa)
                       : "memory");
        pa += stepa;
      }
    }
    asm volatile("cp.async.commit_group;\n cp.async.wait_group 0;" ::: "memory");
  }
  __syncwarp();

  // ---- A4: intensity-centroid angle; lane v+15 reduces row v of the disc with dp4a (masked words x {1, u})
  int m10 = 0, m01 = 0;
  if (lane < kAngleRows) {
    const int o = (cx - kAngleR) & 3;
    const unsigned* urow = g_angle_u[o];
    const unsigned* prow = sa + lane * kAngleRowWords;
    int sum = 0, su = 0;
#pragma unroll
    for (int j = 0; j < kAngleRowWords; ++j) {
      const unsigned wv = prow[j] & g_angle_mask[o][j][lane];
      sum = dp4a_u8_s8(wv, 0x01010101u, sum);
      su = dp4a_u8_s8(wv, urow[j], su);
    }
    m10 = su;
    m01 = (lane - kAngleR) * sum;
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    m10 += __shfl_xor_sync(VSLAM_FULL, m10, o);
    m01 += __shfl_xor_sync(VSLAM_FULL, m01, o);
  }
  const float angle = fast_atan2_deg(__int2float_rn(m01), __int2float_rn(m10));

  if (lane == 0) {
    KeyPoint28 o;
    o.x = __fmul_rn(__int2float_rn(cx), L.scale);
    o.y = __fmul_rn(__int2float_rn(cy), L.scale);
    o.size = __fmul_rn(31.0f, L.scale);
    o.angle = angle;
    o.response = kp.resp;
    o.octave = l;
    o.class_id = -1;
    unsigned char* kpbase = eye ? F.kp[1] : F.kp[0];  // no dynamic indexing of the by-value parameter struct
    reinterpret_cast<KeyPoint28*>(kpbase + (long long)frame * F.kp_stride)[j] = o;
  }
  // ---- A6: angle *= (float)(CV_PI/180.f); a = (float)cos(angle); b = (float)sin(angle)  (double cos/sin of the f32)
  const float rad = __fmul_rn(angle, (float)(3.141592653589793238462643383279502884 / 180.0));
  float a, b;
  sincos_f64_to_f32(rad, a, b);
  // cvRound of a value in (-32, 32): adding 1.5 * 2^23 leaves round-half-even(v) in the low mantissa bits, i.e. the
  // float's bit pattern is 0x4B400000 + round(v).  The bias is not subtracted per coordinate: with row stride 44 the
  // sample address is center + 44 * iy + ix = (center - 45 * 0x4B400000) + 44 * raw_y + raw_x in wrap-around 32-bit
  // arithmetic (shared-window addresses are 32 bits wide).
  constexpr unsigned kRowB = kBriefRowWords * 4;
  const unsigned center = (unsigned)__cvta_generic_to_shared(sb) + kBriefR * kRowB + (unsigned)(cx - xb0) -
                          (kRowB + 1u) * 0x4B400000u;
  const unsigned byte = steered_brief_byte<kRowB>(center, a, b, lane);
  unsigned char* descbase = eye ? F.desc[1] : F.desc[0];
  (descbase + (long long)frame * F.desc_stride)[(size_t)j * VSLAM_DESC_BYTES + lane] = (unsigned char)byte;
}

// ----------------------------------------------------------------------------------------------------------
// A4/A6 with the patches staged by TMA (cp.async.bulk.tensor): k_brief is bound by L1 wavefronts, and almost half of
// them are the two patch copies (68 image rows = 68 cache lines per keypoint through the LSU).  Here lane 0 of a warp
// issues two tensor-tile loads — kTmaBriefW x 37 bytes of the blurred level and 48 x 31 bytes of the un-blurred level —
// that travel L2 -> shared memory without touching L1 or a register, and signal an mbarrier.  The box must start on a
// 16-byte boundary of the row (measured: tools/ubench/tma_probe.cu — an unaligned innermost coordinate of a u8 tensor
// raises an illegal-instruction fault on sm_100a), so the boxes start at x & ~15 and are 15 bytes wider than the
// windows.  The arithmetic is k_brief's, bit for bit.
// Orientation mapping: lane = (word j = lane & 7, row group g = lane >> 3) reduces word j of the 31-px window of rows
// g, g + 4, ...: two shared-memory words and a funnel shift bring the window's word j into place whatever the
// alignment, so one mask table serves every keypoint.
// ----------------------------------------------------------------------------------------------------------
struct BriefMaps {
  CUtensorMap blur[VSLAM_MAX_LEVELS];  // blurred slab, level l: (pitch, h, image)
  CUtensorMap pyr[VSLAM_MAX_LEVELS];   // pyramid slab, levels >= 1: (pitch, h, image)
  CUtensorMap in_l, in_r;              // level 0 of the left / right eye: (pitch, H, frame)
};
#ifndef VSLAM_TMA_BRIEF_W
#define VSLAM_TMA_BRIEF_W 64
#endif
constexpr int kTmaBriefW = VSLAM_TMA_BRIEF_W, kTmaAngleW = 48;  // >= 37 + 15 and 31 + 15, multiples of 16
static_assert(kTmaBriefW % 16 == 0 && kTmaBriefW >= kBriefRows + 15, "BRIEF box: 16-byte granules, window + alignment slack");
constexpr int kTmaBriefBytes = kTmaBriefW * kBriefRows, kTmaAngleBytes = kTmaAngleW * kAngleRows;  // 1776, 992
constexpr int kTmaBriefSlot = (kTmaBriefBytes + 127) & ~127, kTmaAngleSlot = (kTmaAngleBytes + 127) & ~127;  // 1792, 1024

// g_angle_mask_t[k][lane]: byte b = 0xff iff pixel u = 4 j + b - 15 of patch row r = g + 4 k (j = lane & 7, g = lane >> 3)
// lies inside the 31-px disc (0 for r > 30).  Filled by upload_tables.
__device__ unsigned g_angle_mask_t[8][32];

__device__ __forceinline__ void tma_load_3d(unsigned dst, const CUtensorMap* map, int c0, int c1, int c2, unsigned mbar) {
  asm volatile(
      "cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];" ::"r"(dst),
      "l"(map), "r"(c0), "r"(c1), "r"(c2), "r"(mbar)
      : "memory");
}

__global__ void __launch_bounds__(32 * kBriefWarps) k_brief_tma(const __grid_constant__ PyrP P, const __grid_constant__ BriefMaps M,
                                                                 const LvlKp* __restrict__ lvl_kp, const int* __restrict__ lvl_cnt,
                                                                 const FrameLayout F) {
  __shared__ __align__(128) uint8_t s_patch[kBriefWarps][kTmaBriefSlot + kTmaAngleSlot];
  __shared__ __align__(8) unsigned long long s_bar[kBriefWarps];
  const int img = blockIdx.y, lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int j = blockIdx.x * kBriefWarps + warp;
  const int frame = P.eyes == 2 ? img >> 1 : img, eye = P.eyes == 2 ? (img & 1) : 0;
  const int* cnt = lvl_cnt + (long long)img * P.nlevels;
  int total = 0, l = -1, jl = 0;
  {
    int c[VSLAM_MAX_LEVELS];
#pragma unroll
    for (int k = 0; k < VSLAM_MAX_LEVELS; ++k) c[k] = k < P.nlevels ? cnt[k] : 0;
#pragma unroll
    for (int k = 0; k < VSLAM_MAX_LEVELS; ++k) {
      if (l < 0 && j < total + c[k]) {
        l = k;
        jl = j - total;
      }
      total += c[k];
    }
  }
  ResHdr* hdr = F.hdr + frame;
  if (blockIdx.x == 0 && threadIdx.x == 0) {
    hdr->n_kp[eye] = min(total, P.kp_cap);
    if (total > P.kp_cap) atomicMax(&hdr->status, 4 /* VSLAM_ERR_KEYPOINT_OVERFLOW */);
  }
  if (l < 0 || j >= P.kp_cap) return;  // warp-uniform; no block-wide barrier in this kernel
  const LevelP& L = P.lv[l];
  const LvlKp kp = lvl_kp[((long long)img * P.nlevels + l) * P.kp_cap + jl];
  const int cx = (int)(kp.xy & 0xffffu), cy = (int)(kp.xy >> 16);

  const unsigned sb = (unsigned)__cvta_generic_to_shared(s_patch[warp]);
  const unsigned sa = sb + kTmaBriefSlot;
  const unsigned mbar = (unsigned)__cvta_generic_to_shared(&s_bar[warp]);
  if (lane == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(mbar));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(mbar), "r"((unsigned)(kTmaBriefBytes + kTmaAngleBytes))
                 : "memory");
    tma_load_3d(sb, &M.blur[l], (cx - kBriefR) & ~15, cy - kBriefR, img, mbar);
    if (l == 0) tma_load_3d(sa, eye ? &M.in_r : &M.in_l, (cx - kAngleR) & ~15, cy - kAngleR, frame, mbar);
    else tma_load_3d(sa, &M.pyr[l], (cx - kAngleR) & ~15, cy - kAngleR, img, mbar);
  }
  __syncwarp();
  {
    unsigned done = 0;
    while (!done)
      asm volatile(
          "{\n .reg .pred p;\n mbarrier.try_wait.parity.shared::cta.b64 p, [%1], 0;\n selp.u32 %0, 1, 0, p;\n}\n"
          : "=r"(done)
          : "r"(mbar)
          : "memory");
  }

  // ---- A4: intensity-centroid angle.  lane = (j = lane & 7, g = lane >> 3): word j of rows g + 4 k
  int m10 = 0, m01 = 0;
  {
    const int jw = lane & 7, g = lane >> 3;
    unsigned uw = 0;  // (int8) u of the four pixels of word jw
#pragma unroll
    for (int b = 0; b < 4; ++b) uw |= (unsigned)(uint8_t)(int8_t)(4 * jw + b - kAngleR) << (8 * b);
    const unsigned oa = (unsigned)(cx - kAngleR) & 15u;            // byte offset of the window inside the box
    const unsigned wa = sa + (oa & ~3u) + 4u * (unsigned)jw, sha = 8u * (oa & 3u);
    int su = 0, sv = 0;
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      const int r = g + 4 * k;
      if (r < kAngleRows) {
        const unsigned a = wa + (unsigned)(r * kTmaAngleW);
        const unsigned wv = __funnelshift_r(lds32<0>(a), lds32<4>(a), sha) & g_angle_mask_t[k][lane];
        su = dp4a_u8_s8(wv, uw, su);
        sv += (r - kAngleR) * dp4a_u8_s8(wv, 0x01010101u, 0);
      }
    }
    m10 = su;
    m01 = sv;
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    m10 += __shfl_xor_sync(VSLAM_FULL, m10, o);
    m01 += __shfl_xor_sync(VSLAM_FULL, m01, o);
  }
  const float angle = fast_atan2_deg(__int2float_rn(m01), __int2float_rn(m10));

  if (lane == 0) {
    KeyPoint28 o;
    o.x = __fmul_rn(__int2float_rn(cx), L.scale);
    o.y = __fmul_rn(__int2float_rn(cy), L.scale);
    o.size = __fmul_rn(31.0f, L.scale);
    o.angle = angle;
    o.response = kp.resp;
    o.octave = l;
    o.class_id = -1;
    unsigned char* kpbase = eye ? F.kp[1] : F.kp[0];
    reinterpret_cast<KeyPoint28*>(kpbase + (long long)frame * F.kp_stride)[j] = o;
  }
  const float rad = __fmul_rn(angle, (float)(3.141592653589793238462643383279502884 / 180.0));
  float a, b;
  sincos_f64_to_f32(rad, a, b);
  // sample address = patch centre + 48 * iy + ix with the cvRound bias folded in (see k_brief)
  constexpr unsigned kRowB = kTmaBriefW;
  const unsigned center = sb + kBriefR * kRowB + (unsigned)kBriefR + ((unsigned)(cx - kBriefR) & 15u) - (kRowB + 1u) * 0x4B400000u;
  const unsigned byte = steered_brief_byte<kRowB>(center, a, b, lane);
  unsigned char* descbase = eye ? F.desc[1] : F.desc[0];
  (descbase + (long long)frame * F.desc_stride)[(size_t)j * VSLAM_DESC_BYTES + lane] = (unsigned char)byte;
}

}  // namespace vk
