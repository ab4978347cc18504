// `classic` mode: the upstream-ORB-SLAM3 style stages `north_star` names and the reference does
// NOT contain — per-cell FAST with a threshold retry, quadtree keypoint distribution, row-band stereo matching with
// an octave-scaled radius, 11x11 SAD refinement over 11 shifts and a parabolic sub-pixel disparity.
//
// NOT REFERENCE PARITY: jurmy24/orb-slam3-rust has no code for any of this.  The specification is this
// repository's own CPU model, oracle/classic_model.cpp (which follows the published ORB-SLAM3 algorithm and fixes the
// places upstream leaves to implementation accidents); the kernels below are tested bit-for-bit against that model.
// Pyramid, IC angle, blur and steered BRIEF are the parity-mode kernels (orb_kernels.cuh).
//
//   k_cfast_score   S - 1 of every pixel with S > minTh (u8 map per level, written into the not-yet-used blur slab)
//   k_cell_select   one warp per 35-px cell: cell-clipped NMS, corners at iniTh, retry at minTh if the cell has none
//   k_quadtree      one block per (image, level): DistributeOctTree, level-synchronous, deterministic variant
//   k_cstereo       one block per frame: row-band Hamming match (+-1 octave, dist < 75), SAD + parabola, median filter
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "orb_kernels.cuh"
#include "stereo_kernels.cuh"
#include "vslam_params.h"

namespace vk {

constexpr int kCEdge = 19, kCCell = 35;

struct CCell {
  short x0, x1, y0, y1;  // interior [x0, x1) x [y0, y1) in level coordinates (where corners can be)
  int rec_off;           // record offset inside the level's candidate slab
  int cap;               // records the cell can emit (NMS worst case)
};
struct CLevel {
  int cell0, ncells;     // cells of the level inside the per-image cell table
  int rec_off;           // first record of the level inside the per-image candidate slab
  int rec_cap;
  int W, H;              // maxBX - minBX, maxBY - minBY (the quadtree's root extent)
  int nini;              // initial nodes
  float hx;              // W / nini
};
struct ClassicP {
  int ini_th, min_th;
  int total_cells;
  long long rec_img_stride;  // candidate records per image
  CLevel lv[VSLAM_MAX_LEVELS];
};

// ---- S - 1 of every pixel (0 unless S > min_th).  Block = 256 threads = a tile of 64 aligned 4-pixel words x 16 rows of
//      rows [19, h - 19): the tile and its 3-row / 1-word apron are staged in shared memory with coalesced loads, every
//      thread then scores 4 words (rows ty, ty + 4, ty + 8, ty + 12 of column tx).  A word whose four pixels all fail the
//      byte-SIMD necessary test of k_fast's phase A (one of every antipodal pair {k, k + 8}, k = 0, 2, 4, 6, must differ
//      from the centre by more than the threshold) skips the exact DPX scoring: its scores are 0 either way.
constexpr int kCfTW = 64, kCfTH = 16;
__global__ void __launch_bounds__(256) k_cfast_score(const __grid_constant__ PyrP P, int l, int min_th, int tiles_x,
                                                     const uint8_t* __restrict__ in_l, const uint8_t* __restrict__ in_r,
                                                     const uint8_t* __restrict__ pyr, uint8_t* __restrict__ score) {
  __shared__ unsigned s_t[kCfTH + 6][kCfTW + 2];
  const LevelP& L = P.lv[l];
  const int img = blockIdx.y;
  const int x_lo = kCEdge & ~3, nwx = ((L.w - kCEdge + 3) >> 2) - (x_lo >> 2);  // words covering [19, w - 19)
  const int rows = L.h - 2 * kCEdge;
  const int tile_y = blockIdx.x / tiles_x, tile_x = blockIdx.x - tile_y * tiles_x;
  const int wx0 = tile_x * kCfTW, ry0 = tile_y * kCfTH;
  int pitch;
  const uint8_t* base = level_ptr(P, l, img, in_l, in_r, pyr, pitch);
  const int pw = pitch >> 2;  // words per row (pitch is a multiple of 16)
  const uint32_t* gw = reinterpret_cast<const uint32_t*>(base);
  for (int i = threadIdx.x; i < (kCfTH + 6) * (kCfTW + 2); i += 256) {
    const int r = i / (kCfTW + 2), c = i - r * (kCfTW + 2);
    const int gy = min(kCEdge + ry0 + r - 3, L.h - 1);                     // rows past the level only feed unused words
    const int gx = min((x_lo >> 2) + wx0 + c - 1, pw - 1);
    s_t[r][c] = gw[(size_t)gy * pw + gx];
  }
  __syncthreads();
  const int tx = threadIdx.x & 63, ty = threadIdx.x >> 6;
  const int wx = wx0 + tx;
  if (wx >= nwx) return;
  const unsigned t = (unsigned)min_th, t2 = t * 0x00010001u, tadd = (127u - t) * 0x01010101u;
  for (int k = 0; k < 4; ++k) {
    const int rr = ty + 4 * k, ry = ry0 + rr;
    if (ry >= rows) break;
    unsigned w[7][3];
#pragma unroll
    for (int j = 0; j < 7; ++j) {
      w[j][0] = s_t[rr + j][tx];
      w[j][1] = s_t[rr + j][tx + 1];
      w[j][2] = s_t[rr + j][tx + 2];
    }
#define VS_ROW(dy) w[(dy) + 3][0], w[(dy) + 3][1], w[(dy) + 3][2]
    const unsigned cw = w[3][1];
    unsigned out = 0;
    {
      const unsigned a0 = __vabsdiffu4(w[6][1], cw), a8 = __vabsdiffu4(w[0][1], cw);
      const unsigned a4 = __vabsdiffu4(shifted4<3>(VS_ROW(0)), cw), a12 = __vabsdiffu4(shifted4<-3>(VS_ROW(0)), cw);
      const unsigned a2 = __vabsdiffu4(shifted4<2>(VS_ROW(2)), cw), a14 = __vabsdiffu4(shifted4<-2>(VS_ROW(2)), cw);
      const unsigned a6 = __vabsdiffu4(shifted4<2>(VS_ROW(-2)), cw), a10 = __vabsdiffu4(shifted4<-2>(VS_ROW(-2)), cw);
      // per byte "v > t" as bit 7 of (v + 127 - t) | v (a carry can only add false positives of this necessary test)
      const unsigned pass = ((a0 + tadd) | a0 | (a8 + tadd) | a8) & ((a4 + tadd) | a4 | (a12 + tadd) | a12) &
                            ((a2 + tadd) | a2 | (a10 + tadd) | a10) & ((a6 + tadd) | a6 | (a14 + tadd) | a14) & 0x80808080u;
      if (pass) {
        unsigned sv[16];
        sv[0] = shifted4<0>(VS_ROW(3));
        sv[1] = shifted4<1>(VS_ROW(3));
        sv[2] = shifted4<2>(VS_ROW(2));
        sv[3] = shifted4<3>(VS_ROW(1));
        sv[4] = shifted4<3>(VS_ROW(0));
        sv[5] = shifted4<3>(VS_ROW(-1));
        sv[6] = shifted4<2>(VS_ROW(-2));
        sv[7] = shifted4<1>(VS_ROW(-3));
        sv[8] = shifted4<0>(VS_ROW(-3));
        sv[9] = shifted4<-1>(VS_ROW(-3));
        sv[10] = shifted4<-2>(VS_ROW(-2));
        sv[11] = shifted4<-3>(VS_ROW(-1));
        sv[12] = shifted4<-3>(VS_ROW(0));
        sv[13] = shifted4<-3>(VS_ROW(1));
        sv[14] = shifted4<-2>(VS_ROW(2));
        sv[15] = shifted4<-1>(VS_ROW(3));
#pragma unroll
        for (int half = 0; half < 2; ++half) {
          if (!(pass & (half ? 0x80800000u : 0x00008080u))) continue;
          const unsigned sel = half ? 0x4342u : 0x4140u;
          unsigned p[16];
#pragma unroll
          for (int q = 0; q < 16; ++q) p[q] = __byte_perm(sv[q], 0u, sel);
          const unsigned xs = fast_score2(p, __byte_perm(cw, 0u, sel), t2);  // max(S - t, 0) per 16-bit lane
          const unsigned lo = xs & 0xffffu, hi = xs >> 16;                    // S - 1 = (S - t) + t - 1 where S > t
          const unsigned b0 = lo ? lo + t - 1u : 0u, b1 = hi ? hi + t - 1u : 0u;
          out |= (b0 | (b1 << 8)) << (16 * half);
        }
      }
    }
#undef VS_ROW
    // pixels left of 19 / right of w - 19 inside the first / last word are never read by k_cell_select
    *reinterpret_cast<uint32_t*>(score + (long long)img * P.blur_img_stride + L.blur_off + (size_t)(kCEdge + ry) * L.pitch +
                                 x_lo + 4 * wx) = out;
  }
}

// ---- per-cell selection: one warp per cell.  Candidates of a cell in raster order into its slot.
//      The cell's scores are first copied into a zero-padded shared-memory tile with independent word loads (all in
//      flight at once) — exactly the score buffer cv::FAST would hold for the cell image: everything outside the cell's
//      interior reads as 0, so the NMS needs no border tests.  A pass then scans the tile's words, queues the pixels that
//      pass the threshold in raster order (4 ballots per 128 pixels) and runs the 8-neighbour test on full warps, one
//      queued pixel per lane.
constexpr int kCellTW = 13;    // tile words per row: <= 41 interior pixels + alignment slack + zero ring
constexpr int kCellTH = 46;    // tile rows: <= 41 interior rows + zero ring
constexpr int kCellQ = 448;    // queue entries per warp: the NMS worst case of a 41 x 41 cell is 441
__global__ void __launch_bounds__(128) k_cell_select(const __grid_constant__ PyrP P, const __grid_constant__ ClassicP CP,
                                                     const CCell* __restrict__ cells, const uint8_t* __restrict__ score,
                                                     int* __restrict__ cell_cnt, uint32_t* __restrict__ cell_rec) {
  __shared__ uint32_t s_tile[4][kCellTH * kCellTW];
  __shared__ uint32_t s_q[4][kCellQ];
  const int img = blockIdx.y, lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int ci = blockIdx.x * 4 + warp;
  if (ci >= CP.total_cells) return;
  int l = 0;
  for (int k = 1; k < P.nlevels; ++k)
    if (ci >= CP.lv[k].cell0) l = k;
  const LevelP& L = P.lv[l];
  const CCell c = cells[ci];
  const uint8_t* sm = score + (long long)img * P.blur_img_stride + L.blur_off;
  const int pitch = L.pitch;
  uint32_t* out = cell_rec + (long long)img * CP.rec_img_stride + CP.lv[l].rec_off + c.rec_off;
  uint32_t* q = s_q[warp];
  uint32_t* tile = s_tile[warp];
  const unsigned lt = (1u << lane) - 1u;
  // tile geometry: tile word (tr, tw) holds pixels x = xa + 4 (tw - 1) .. + 3 of row y = c.y0 + tr - 1; column 0 / the
  // last column and row 0 / the last row are the zero ring
  const int xa = c.x0 & ~3, nw = ((c.x1 - 1) >> 2) - (xa >> 2) + 1, nr = c.y1 - c.y0;  // interior words per row, rows
  const bool fits = nw + 2 <= kCellTW && nr + 2 <= kCellTH;
  int n = 0;
  if (fits) {
    for (int i = lane; i < (nr + 2) * kCellTW; i += 32) tile[i] = 0u;
    __syncwarp();
    for (int i = lane; i < nr * nw; i += 32) {
      const int r = i / nw, w = i - r * nw;
      const int x = xa + 4 * w;
      unsigned v = *reinterpret_cast<const uint32_t*>(sm + (size_t)(c.y0 + r) * pitch + x);
      unsigned keepm = 0;  // bytes inside [c.x0, c.x1)
#pragma unroll
      for (int b = 0; b < 4; ++b)
        if (x + b >= c.x0 && x + b < c.x1) keepm |= 0xffu << (8 * b);
      tile[(r + 1) * kCellTW + w + 1] = v & keepm;
    }
    __syncwarp();
    const unsigned char* tb = reinterpret_cast<const unsigned char*>(tile);
    for (int pass = 0; pass < 2 && n == 0; ++pass) {
      const unsigned th = (unsigned)(pass == 0 ? CP.ini_th : CP.min_th);
      int qn = 0;
      for (int i0 = 0; i0 < nr * nw; i0 += 32) {  // scan: words in raster order, lane = word
        const int i = i0 + lane;
        unsigned v = 0;
        int r = 0, w = 0;
        if (i < nr * nw) {
          r = i / nw;
          w = i - r * nw;
          v = tile[(r + 1) * kCellTW + w + 1];
        }
        unsigned cnt_before = 0, mine = 0;
        unsigned mb[4];
#pragma unroll
        for (int b = 0; b < 4; ++b) {
          const unsigned sc = (v >> (8 * b)) & 0xffu;
          const bool cand = sc != 0u && sc + 1u > th;
          mb[b] = __ballot_sync(VSLAM_FULL, cand);
          cnt_before += __popc(mb[b] & lt);
          mine |= cand ? (1u << b) : 0u;
        }
        // raster order inside the warp's 128 pixels: lane-major, then byte
        int pos = qn + (int)cnt_before;
#pragma unroll
        for (int b = 0; b < 4; ++b)
          if (mine & (1u << b)) {
            if (pos < kCellQ) q[pos] = (((v >> (8 * b)) & 0xffu) << 24) | ((unsigned)(c.y0 + r) << 12) | (unsigned)(xa + 4 * w + b);
            ++pos;
          }
        qn += __popc(mb[0]) + __popc(mb[1]) + __popc(mb[2]) + __popc(mb[3]);
      }
      __syncwarp();
      qn = min(qn, kCellQ);  // (cannot exceed the NMS worst case... of corners, but pre-NMS pixels can: see the fallback)
      for (int b0 = 0; b0 < qn; b0 += 32) {  // NMS, one queued pixel per lane
        bool keep = false;
        uint32_t e = 0;
        if (b0 + lane < qn) {
          e = q[b0 + lane];
          const int x = (int)(e & 0xfffu), y = (int)((e >> 12) & 0xfffu), sc = (int)(e >> 24);
          const unsigned char* p = tb + ((y - c.y0 + 1) * kCellTW + 1) * 4 + (x - xa);
          constexpr int RS = kCellTW * 4;
          keep = sc > p[-1] && sc > p[1] && sc > p[-RS - 1] && sc > p[-RS] && sc > p[-RS + 1] && sc > p[RS - 1] && sc > p[RS] &&
                 sc > p[RS + 1];
        }
        const unsigned m = __ballot_sync(VSLAM_FULL, keep);
        if (keep) {
          const int pos = n + __popc(m & lt);
          if (pos < c.cap) out[pos] = e;
        }
        n += __popc(m);
      }
    }
  }
  // Pre-NMS pixels can outnumber the queue only on pathological inputs (> 448 pixels above the threshold in one cell);
  // such a cell — and any cell larger than the tile — takes the direct path: every pixel tested in place.
  bool direct = !fits;
  if (fits) {
    int above = 0;  // pixels above min_th (the larger set) in the tile
    for (int i = lane; i < nr * nw; i += 32) {
      const unsigned v = tile[(i / nw + 1) * kCellTW + (i % nw) + 1];
#pragma unroll
      for (int b = 0; b < 4; ++b) above += ((v >> (8 * b)) & 0xffu) != 0u;
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) above += __shfl_xor_sync(VSLAM_FULL, above, o);
    direct = above > kCellQ;
  }
  if (direct) {
    n = 0;
    for (int pass = 0; pass < 2 && n == 0; ++pass) {
      const int th = pass == 0 ? CP.ini_th : CP.min_th;
      for (int y = c.y0; y < c.y1; ++y)
        for (int xb = c.x0; xb < c.x1; xb += 32) {
          const int x = xb + lane;
          bool keep = false;
          int sc = 0;
          if (x < c.x1) {
            const uint8_t* p = sm + (size_t)y * pitch + x;
            sc = p[0];
            if (sc + 1 > th && sc > 0) {
              const bool xl = x > c.x0, xr = x + 1 < c.x1, yu = y > c.y0, yd = y + 1 < c.y1;
              const int a = xl ? p[-1] : 0, r = xr ? p[1] : 0;
              const int ul = (yu && xl) ? p[-pitch - 1] : 0, uc = yu ? p[-pitch] : 0, ur = (yu && xr) ? p[-pitch + 1] : 0;
              const int dl = (yd && xl) ? p[pitch - 1] : 0, dc = yd ? p[pitch] : 0, dr = (yd && xr) ? p[pitch + 1] : 0;
              keep = sc > a && sc > r && sc > ul && sc > uc && sc > ur && sc > dl && sc > dc && sc > dr;
            }
          }
          const unsigned m = __ballot_sync(VSLAM_FULL, keep);
          if (keep) {
            const int pos = n + __popc(m & lt);
            if (pos < c.cap) out[pos] = ((unsigned)sc << 24) | ((unsigned)y << 12) | (unsigned)x;
          }
          n += __popc(m);
        }
    }
  }
  if (lane == 0) cell_cnt[(long long)img * CP.total_cells + ci] = min(n, c.cap);
}

// ---- DistributeOctTree (deterministic variant of oracle/classic_model.cpp), one block of 256 threads per (image, level).
// Candidate k of the level: key[k] = resp << 32 | y << 16 | x (level coordinates), node_of[k] = index of its live node.
// Live nodes are kept compacted in creation order: after every pass [unsplit nodes in order] ++ [children in split order].
struct QNode {
  short ulx, uly, brx, bry;
};
// dynamic shared memory for `ncap` live nodes (ncap >= max N_l + 8: a pass never holds more than N_l + 3 nodes)
__host__ inline size_t quadtree_smem_bytes(int ncap) { return (size_t)ncap * (2 * 8 + 2 * 4 + 16 + 2 + 2 + 1) + 64; }
__global__ void __launch_bounds__(256) k_quadtree(const __grid_constant__ PyrP P, const __grid_constant__ ClassicP CP, int ncap,
                                                  const CCell* __restrict__ cells, const int* __restrict__ cell_cnt,
                                                  const uint32_t* __restrict__ cell_rec, unsigned long long* __restrict__ keys,
                                                  uint32_t* __restrict__ node_of_g, LvlKp* __restrict__ lvl_kp,
                                                  int* __restrict__ lvl_cnt, const FrameLayout F) {
  extern __shared__ __align__(16) unsigned char q_smem[];
  const int kCMaxNodes = ncap;
  typedef unsigned (*Child4)[4];
  Child4 s_child = reinterpret_cast<Child4>(q_smem);                                // [ncap][4] child key counts / new indices
  QNode* s_node[2];
  s_node[0] = reinterpret_cast<QNode*>(q_smem + (size_t)ncap * 16);                  // [2][ncap]
  s_node[1] = s_node[0] + ncap;
  unsigned* s_cnt[2];
  s_cnt[0] = reinterpret_cast<unsigned*>(s_node[1] + ncap);                          // [2][ncap] keys per node
  s_cnt[1] = s_cnt[0] + ncap;
  unsigned short* s_remap = reinterpret_cast<unsigned short*>(s_cnt[1] + ncap);      // old node -> new index (unsplit)
  unsigned short* s_order = s_remap + ncap;                                          // size-ordered phase, largest first
  unsigned char* s_split = reinterpret_cast<unsigned char*>(s_order + ncap);
  __shared__ int s_scan[256];
  __shared__ int s_i[8];
  const int img = blockIdx.x, l = blockIdx.y, tid = threadIdx.x;
  const CLevel& CL = CP.lv[l];
  const int N = P.lv[l].nfeat;
  const long long base = (long long)img * CP.rec_img_stride + CL.rec_off;
  unsigned long long* key = keys + base;
  uint32_t* node_of = node_of_g + base;
  const int minB = kCEdge - 3;

  // ---- gather the cells' candidates in cell order (exclusive scan of the cell counts, 256 cells per round)
  int n = 0;
  for (int c0 = 0; c0 < CL.ncells; c0 += 256) {
    const int ci = c0 + tid;
    const int cnt = ci < CL.ncells ? cell_cnt[(long long)img * CP.total_cells + CL.cell0 + ci] : 0;
    s_scan[tid] = cnt;
    __syncthreads();
    for (int d = 1; d < 256; d <<= 1) {
      const int v = tid >= d ? s_scan[tid - d] : 0;
      __syncthreads();
      s_scan[tid] += v;
      __syncthreads();
    }
    const int off = n + s_scan[tid] - cnt;
    if (ci < CL.ncells) {
      const uint32_t* src = cell_rec + base + cells[CL.cell0 + ci].rec_off;
      for (int i = 0; i < cnt; ++i) {
        const uint32_t r = src[i];
        key[off + i] = ((unsigned long long)(r >> 24) << 32) | (((r >> 12) & 0xfffu) << 16) | (r & 0xfffu);
      }
    }
    n += s_scan[255];
    __syncthreads();
  }
  int cur = 0, live = 0;
  if (n > 0 && N > 0 && N + 8 <= ncap) {
    // ---- initial nodes
    for (int i = tid; i < CL.nini; i += 256) {
      s_node[0][i] = QNode{(short)(int)(CL.hx * (float)i), 0, (short)(int)(CL.hx * (float)(i + 1)), (short)CL.H};
      s_cnt[0][i] = 0;
    }
    __syncthreads();
    for (int k = tid; k < n; k += 256) {
      const int x = (int)(key[k] & 0xffffu) - minB;
      const int i = min((int)(__fdiv_rn((float)x, CL.hx)), CL.nini - 1);
      node_of[k] = (uint32_t)i;
      atomicAdd(&s_cnt[0][i], 1u);
    }
    __syncthreads();
    if (tid == 0) {  // drop empty initial nodes (nini is 1..4)
      int m = 0;
      for (int i = 0; i < CL.nini; ++i) {
        s_remap[i] = (unsigned short)m;
        if (s_cnt[0][i] > 0) {
          s_node[1][m] = s_node[0][i];
          s_cnt[1][m] = s_cnt[0][i];
          ++m;
        }
      }
      s_i[0] = m;
    }
    __syncthreads();
    live = s_i[0];
    for (int k = tid; k < n; k += 256) node_of[k] = s_remap[node_of[k]];
    cur = 1;
    __syncthreads();

    bool careful = false, finish = false;
    while (!finish) {
      const QNode* nd = s_node[cur];
      const unsigned* cnt = s_cnt[cur];
      // child key counts of every expandable node
      for (int i = tid; i < live; i += 256) {
        s_child[i][0] = s_child[i][1] = s_child[i][2] = s_child[i][3] = 0;
        s_split[i] = 0;
      }
      __syncthreads();
      for (int k = tid; k < n; k += 256) {
        const int i = (int)node_of[k];
        if (cnt[i] > 1) {
          const QNode q = nd[i];
          const int mx = q.ulx + ((q.brx - q.ulx + 1) >> 1), my = q.uly + ((q.bry - q.uly + 1) >> 1);  // ceil(d / 2)
          const int x = (int)(key[k] & 0xffffu) - minB, y = (int)((key[k] >> 16) & 0xffffu) - minB;
          atomicAdd(&s_child[i][(x < mx ? 0 : 1) + (y < my ? 0 : 2)], 1u);
        }
      }
      __syncthreads();
      // which nodes split in this pass
      if (!careful) {
        for (int i = tid; i < live; i += 256) s_split[i] = cnt[i] > 1;
        __syncthreads();
      } else {
        // size order: count descending, creation index descending (rank by counting)
        for (int i = tid; i < live; i += 256) {
          if (cnt[i] > 1) {
            int rank = 0;
            for (int j = 0; j < live; ++j)
              if (cnt[j] > 1 && (cnt[j] > cnt[i] || (cnt[j] == cnt[i] && j > i))) ++rank;
            s_order[rank] = (unsigned short)i;
          }
        }
        __syncthreads();
        if (tid == 0) {
          int ne = 0;
          for (int i = 0; i < live; ++i) ne += cnt[i] > 1;
          int lv = live;
          for (int r = 0; r < ne; ++r) {
            const int i = s_order[r];
            s_split[i] = 1;
            lv += (int)(s_child[i][0] > 0) + (int)(s_child[i][1] > 0) + (int)(s_child[i][2] > 0) + (int)(s_child[i][3] > 0) - 1;
            if (lv >= N) break;
          }
          s_i[1] = ne;
        }
        __syncthreads();
      }
      // new node list: unsplit nodes in order, then the children in split order (creation order n1..n4)
      if (tid == 0) {
        QNode* nn = s_node[cur ^ 1];
        unsigned* nc = s_cnt[cur ^ 1];
        int m = 0;
        for (int i = 0; i < live; ++i)
          if (!s_split[i]) {
            s_remap[i] = (unsigned short)m;
            nn[m] = nd[i];
            nc[m] = cnt[i];
            ++m;
          }
        int nexp = 0, overflow = 0;
        auto emit_children = [&](int i) {
          const QNode q = nd[i];
          const short mx = (short)(q.ulx + ((q.brx - q.ulx + 1) >> 1)), my = (short)(q.uly + ((q.bry - q.uly + 1) >> 1));
          const QNode ch[4] = {{q.ulx, q.uly, mx, my}, {mx, q.uly, q.brx, my}, {q.ulx, my, mx, q.bry}, {mx, my, q.brx, q.bry}};
          for (int c = 0; c < 4; ++c) {
            const unsigned k = s_child[i][c];
            if (k > 0) {
              if (m >= kCMaxNodes) {
                overflow = 1;
                s_child[i][c] = 0;
                continue;
              }
              nn[m] = ch[c];
              nc[m] = k;
              nexp += k > 1;
              s_child[i][c] = (unsigned)m;
              ++m;
            }
          }
        };
        if (!careful) {
          for (int i = 0; i < live; ++i)
            if (s_split[i]) emit_children(i);
        } else {
          for (int r = 0; r < s_i[1]; ++r)
            if (s_split[s_order[r]]) emit_children(s_order[r]);
        }
        s_i[2] = m;
        s_i[3] = nexp;
        s_i[4] = overflow;
      }
      __syncthreads();
      const int prev = live;
      for (int k = tid; k < n; k += 256) {
        const int i = (int)node_of[k];
        if (s_split[i]) {
          const QNode q = nd[i];
          const int mx = q.ulx + ((q.brx - q.ulx + 1) >> 1), my = q.uly + ((q.bry - q.uly + 1) >> 1);
          const int x = (int)(key[k] & 0xffffu) - minB, y = (int)((key[k] >> 16) & 0xffffu) - minB;
          node_of[k] = s_child[i][(x < mx ? 0 : 1) + (y < my ? 0 : 2)];
        } else {
          node_of[k] = s_remap[i];
        }
      }
      live = s_i[2];
      const int nexp = s_i[3];
      if (s_i[4]) atomicMax(&F.hdr[img / P.eyes].status, 4 /* node table overflow: reported like a keypoint overflow */);
      cur ^= 1;
      __syncthreads();
      if (live >= N || live == prev || s_i[4]) finish = true;
      else if (!careful && live + nexp * 3 > N) careful = true;
    }
  }
  // ---- one keypoint per node: maximum response, first in candidate order
  unsigned long long* s_best = reinterpret_cast<unsigned long long*>(&s_child[0][0]);  // [live] (>= 16 B per node there)
  for (int i = tid; i < live; i += 256) s_best[i] = 0ull;
  __syncthreads();
  for (int k = tid; k < n; k += 256) {
    const unsigned long long v = ((key[k] >> 32) << 32) | (unsigned long long)(0xffffffffu - (unsigned)k);
    atomicMax(&s_best[node_of[k]], v);
  }
  __syncthreads();
  int m2 = min(live, P.kp_cap);
  if (live > P.kp_cap && tid == 0) atomicMax(&F.hdr[img / P.eyes].status, 4);
  LvlKp* out = lvl_kp + ((long long)img * P.nlevels + l) * P.kp_cap;
  for (int i = tid; i < live; i += 256) {  // emit in candidate order (rank by counting)
    const unsigned ki = 0xffffffffu - (unsigned)(s_best[i] & 0xffffffffu);
    int rank = 0;
    for (int j = 0; j < live; ++j) rank += (0xffffffffu - (unsigned)(s_best[j] & 0xffffffffu)) < ki;
    if (rank < m2) {
      const unsigned long long kv = key[ki];
      LvlKp o;
      o.xy = (unsigned)(((kv >> 16) & 0xffffu) << 16) | (unsigned)(kv & 0xffffu);
      o.resp = (float)(unsigned)(kv >> 32);
      out[rank] = o;
    }
  }
  if (tid == 0) lvl_cnt[(long long)img * P.nlevels + l] = m2;
}

// ---- classic stereo: one block of 1024 threads per frame.
struct CStereoP {
  float mbf, mb;       // (float)(fx * baseline), (float)baseline
  double fx, fy, cx, cy;
};
__host__ inline size_t cstereo_smem_bytes(int cap, int H) {
  return (size_t)cap * (8 + 36 + 4 + 4 + 4) + (size_t)((cap + 1) & ~1) * 2 + (size_t)(H + 2) * 4 * 2 + 64;
}
__global__ void __launch_bounds__(1024) k_cstereo(const __grid_constant__ PyrP P, const FrameLayout F, const CStereoP S, int H, int band,
                                                  const uint8_t* __restrict__ in_l, const uint8_t* __restrict__ in_r,
                                                  const uint8_t* __restrict__ pyr, float* __restrict__ g_uright,
                                                  float* __restrict__ g_depth) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  __shared__ int s_part[32];
  __shared__ int s_misc[4];
  const int cap = F.cap;
  float2* s_xy = reinterpret_cast<float2*>(smem_raw);                 // [cap] right keypoints sorted by row
  unsigned* s_desc = reinterpret_cast<unsigned*>(s_xy + cap);         // [cap][9]
  int* s_match = reinterpret_cast<int*>(s_desc + (size_t)cap * 9);    // [cap] per left: right index or -1
  int* s_odist = s_match + cap;                                       // [cap] ORB distance
  int* s_sad = s_odist + cap;                                         // [cap] SAD of the accepted match or -1
  int* s_start = s_sad + cap;                                         // [H + 1] row bucket offsets
  int* s_fill = s_start + H + 2;                                      // [H + 1]
  unsigned short* s_idx = reinterpret_cast<unsigned short*>(s_fill + H + 2);  // [cap] original right index
  const int frame = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  ResHdr* hdr = F.hdr + frame;
  const int nL = min(hdr->n_kp[0], cap), nR = min(hdr->n_kp[1], cap);
  const unsigned TH_HIGH = 100u, TH_ORB = 75u;  // thOrbDist = (TH_HIGH + TH_LOW) / 2

  // ---- counting sort of the right keypoints by floor(y)
  for (int i = tid; i <= H; i += 1024) s_start[i] = 0;
  __syncthreads();
  for (int r = tid; r < nR; r += 1024) {
    const int b = min(max((int)floorf(kp_ptr(F, frame, 1, r)[1]), 0), H - 1);
    atomicAdd(&s_start[b], 1);
  }
  __syncthreads();
  if (warp == 0) {  // exclusive scan over H buckets by one warp
    int run = 0;
    for (int b0 = 0; b0 < H; b0 += 32) {
      const int i = b0 + lane;
      const int v = i < H ? s_start[i] : 0;
      int incl = v;
#pragma unroll
      for (int d = 1; d < 32; d <<= 1) {
        const int o = __shfl_up_sync(VSLAM_FULL, incl, d);
        if (lane >= d) incl += o;
      }
      if (i < H) {
        s_start[i] = run + incl - v;
        s_fill[i] = run + incl - v;
      }
      run += __shfl_sync(VSLAM_FULL, incl, 31);
    }
    if (lane == 0) s_start[H] = run;
  }
  __syncthreads();
  const unsigned* descR = reinterpret_cast<const unsigned*>(F.desc[1] + (long long)frame * F.desc_stride);
  for (int r = tid; r < nR; r += 1024) {
    const float* k = kp_ptr(F, frame, 1, r);
    const int b = min(max((int)floorf(k[1]), 0), H - 1);
    const int p = atomicAdd(&s_fill[b], 1);
    s_xy[p] = make_float2(k[0], k[1]);
    s_idx[p] = (unsigned short)r;
    const uint4 d0 = *reinterpret_cast<const uint4*>(descR + (size_t)r * 8);
    const uint4 d1 = *reinterpret_cast<const uint4*>(descR + (size_t)r * 8 + 4);
    unsigned* dp = s_desc + (size_t)p * 9;
    dp[0] = d0.x; dp[1] = d0.y; dp[2] = d0.z; dp[3] = d0.w;
    dp[4] = d1.x; dp[5] = d1.y; dp[6] = d1.z; dp[7] = d1.w;
    dp[8] = (unsigned)reinterpret_cast<const int*>(k)[5];  // octave
  }
  __syncthreads();

  // ---- row-band match: 8 lanes per left keypoint.  Right keypoint iR is a candidate of image row `row` iff
  //      floor(yR - r) <= row <= ceil(yR + r), r = 2 * scale[octave_R]; its bucket floor(yR) is then within +-band rows
  //      (band = ceil(2 * largest level scale) + 2, from the host).
  const float maxD = __fdiv_rn(S.mbf, S.mb);
  const unsigned* descL = reinterpret_cast<const unsigned*>(F.desc[0] + (long long)frame * F.desc_stride);
  const int sub = tid & 7;
  for (int l0 = 0; l0 < nL; l0 += 128) {
    const int li = l0 + (tid >> 3);
    const bool valid = li < nL;
    unsigned key = (TH_HIGH << 16) | 0xffffu;
    if (valid) {
      const float* kl = kp_ptr(F, frame, 0, li);
      const float uL = kl[0], vL = kl[1];
      const int levelL = reinterpret_cast<const int*>(kl)[5];
      const int row = (int)vL;
      const float minU = __fsub_rn(uL, maxD), maxU = uL;
      if (row >= 0 && row < H && !(maxU < 0.f)) {
        const uint4 a0 = *reinterpret_cast<const uint4*>(descL + (size_t)li * 8);
        const uint4 a1 = *reinterpret_cast<const uint4*>(descL + (size_t)li * 8 + 4);
        const int b0 = max(row - band, 0), b1 = min(row + band, H - 1);
        const int p1 = s_start[b1 + 1];
        for (int p = s_start[b0] + sub; p < p1; p += 8) {
          const float2 q = s_xy[p];
          const unsigned* dr = s_desc + (size_t)p * 9;
          const int octR = (int)dr[8];
          const float r = __fmul_rn(2.0f, P.lv[octR].scale);
          if (row < (int)floorf(__fsub_rn(q.y, r)) || row > (int)ceilf(__fadd_rn(q.y, r))) continue;
          if (octR < levelL - 1 || octR > levelL + 1) continue;
          if (!(q.x >= minU && q.x <= maxU)) continue;
          const unsigned d = __popc(a0.x ^ dr[0]) + __popc(a0.y ^ dr[1]) + __popc(a0.z ^ dr[2]) + __popc(a0.w ^ dr[3]) +
                             __popc(a1.x ^ dr[4]) + __popc(a1.y ^ dr[5]) + __popc(a1.z ^ dr[6]) + __popc(a1.w ^ dr[7]);
          const unsigned k2 = (d << 16) | (unsigned)s_idx[p];
          key = min(key, k2);  // smallest distance, then smallest right index (== first in candidate order)
        }
      }
    }
    key = __reduce_min_sync(0xffu << (lane & 24), key);  // the 8 lanes that share this left keypoint
    if (valid && sub == 0) {
      const unsigned best = key >> 16;
      s_match[li] = best < TH_ORB ? (int)(key & 0xffffu) : -1;
      s_odist[li] = (int)best;
      s_sad[li] = -1;
    }
  }
  __syncthreads();

  // ---- SAD refinement + parabola: one warp per accepted left keypoint
  for (int li = warp; li < nL; li += 32) {
    const int iR = s_match[li];
    if (iR < 0) continue;  // warp-uniform
    const float* kl = kp_ptr(F, frame, 0, li);
    const float* kr = kp_ptr(F, frame, 1, iR);
    const float uL = kl[0], vL = kl[1], uR0 = kr[0];
    const int lev = reinterpret_cast<const int*>(kl)[5];
    const LevelP& LV = P.lv[lev];
    const float inv = __fdiv_rn(1.0f, LV.scale);
    const int su = (int)roundf(__fmul_rn(uL, inv)), sv = (int)roundf(__fmul_rn(vL, inv)), sr0 = (int)roundf(__fmul_rn(uR0, inv));
    const int w = 5, Ls = 5;
    bool ok = !(sv - w < 0 || sv + w >= LV.h || su - w < 0 || su + w >= LV.w);
    const int iniu = sr0 + Ls - w, endu = sr0 + Ls + w + 1;
    ok = ok && !(iniu < 0 || endu >= LV.w) && !(sr0 - Ls - w < 0);
    int bestsad = 0x7fffffff, bestinc = 0, sads_m1 = 0, sads_0 = 0, sads_p1 = 0;
    if (ok) {
      int pl, pr;
      const uint8_t* IL = level_ptr(P, lev, 2 * frame, in_l, in_r, pyr, pl);
      const uint8_t* IR = level_ptr(P, lev, 2 * frame + 1, in_l, in_r, pyr, pr);
      // lane t < 121 owns window pixel (dy, dx) = (t / 11 - 5, t % 11 - 5) ... 4 lanes-worth: 121 = 3 * 32 + 25
      int lv4[4], off4[4];
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        const int t = lane + 32 * k;
        const int dy = t / 11 - w, dx = t - (t / 11) * 11 - w;
        off4[k] = t < 121 ? (sv + dy) * pr + dx : -1;
        lv4[k] = t < 121 ? (int)IL[(size_t)(sv + dy) * pl + su + dx] : 0;
      }
      int sad[11];
#pragma unroll
      for (int inc = -5; inc <= 5; ++inc) {
        int s = 0;
#pragma unroll
        for (int k = 0; k < 4; ++k)
          if (off4[k] >= 0) s += abs(lv4[k] - (int)IR[off4[k] + sr0 + inc]);
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(VSLAM_FULL, s, o);
        sad[inc + 5] = s;
      }
#pragma unroll
      for (int i = 0; i < 11; ++i)
        if (sad[i] < bestsad) {
          bestsad = sad[i];
          bestinc = i - 5;
        }
#pragma unroll
      for (int i = 1; i < 10; ++i)
        if (i - 5 == bestinc) {
          sads_m1 = sad[i - 1];
          sads_0 = sad[i];
          sads_p1 = sad[i + 1];
        }
      ok = !(bestinc == -Ls || bestinc == Ls);
    }
    float bestuR = -1.f, depth = -1.f;
    if (ok) {
      const float d1 = (float)sads_m1, d2 = (float)sads_0, d3 = (float)sads_p1;
      const float den = __fmul_rn(2.0f, __fsub_rn(__fadd_rn(d1, d3), __fmul_rn(2.0f, d2)));
      ok = den != 0.0f;
      if (ok) {
        const float deltaR = __fdiv_rn(__fsub_rn(d1, d3), den);
        ok = !(deltaR < -1.f || deltaR > 1.f);
        if (ok) {
          bestuR = __fmul_rn(LV.scale, __fadd_rn(__fadd_rn((float)sr0, (float)bestinc), deltaR));
          float disparity = __fsub_rn(uL, bestuR);
          ok = disparity >= 0.f && disparity < maxD;
          if (ok) {
            if (disparity <= 0.f) {
              disparity = 0.01f;
              bestuR = __fsub_rn(uL, 0.01f);
            }
            depth = __fdiv_rn(S.mbf, disparity);
          }
        }
      }
    }
    if (lane == 0) {
      if (ok) s_sad[li] = bestsad;
      else s_match[li] = -1;
      g_uright[(long long)frame * cap + li] = ok ? bestuR : -1.f;
      g_depth[(long long)frame * cap + li] = ok ? depth : -1.f;
    }
  }
  for (int li = tid; li < nL; li += 1024)
    if (s_match[li] < 0) {
      g_uright[(long long)frame * cap + li] = -1.f;
      g_depth[(long long)frame * cap + li] = -1.f;
    }
  __syncthreads();

  // ---- median filter: the (n / 2)-th smallest SAD among the accepted matches
  int nv = 0;
  for (int b = 0; b < nL; b += 1024) {
    const int li = b + tid;
    nv += __syncthreads_count(li < nL && s_match[li] >= 0);
  }
  if (tid == 0) s_misc[0] = -1;
  __syncthreads();
  if (nv > 0) {
    const int kth = nv / 2;
    for (int li = tid; li < nL; li += 1024) {
      if (s_match[li] < 0) continue;
      const int v = s_sad[li];
      int less = 0, leq = 0;
      for (int j = 0; j < nL; ++j)
        if (s_match[j] >= 0) {
          less += s_sad[j] < v;
          leq += s_sad[j] <= v;
        }
      if (less <= kth && kth < leq) s_misc[0] = v;  // every thread that hits writes the same value
    }
    __syncthreads();
    const float thDist = __fmul_rn(__fmul_rn(1.5f, 1.4f), (float)s_misc[0]);
    for (int li = tid; li < nL; li += 1024)
      if (s_match[li] >= 0 && !((float)s_sad[li] < thDist)) {
        s_match[li] = -1;
        g_uright[(long long)frame * cap + li] = -1.f;
        g_depth[(long long)frame * cap + li] = -1.f;
      }
  }
  __syncthreads();

  // ---- outputs: Vector<DMatch> in ascending left index, points from the refined depth
  if (tid == 0) s_misc[1] = 0;
  __syncthreads();
  int4* mout = reinterpret_cast<int4*>(F.matches + (long long)frame * F.matches_stride);
  double* pout = reinterpret_cast<double*>(F.pts + (long long)frame * F.pts_stride);
  unsigned char* vout = F.valid + (long long)frame * F.valid_stride;
  for (int b = 0; b < nL; b += 1024) {
    const int li = b + tid;
    const bool has = li < nL && s_match[li] >= 0;
    const unsigned m = __ballot_sync(VSLAM_FULL, has);
    if (lane == 0) s_part[warp] = __popc(m);
    __syncthreads();
    int before = s_misc[1];
    for (int w2 = 0; w2 < warp; ++w2) before += s_part[w2];
    if (has) {
      const int pos = before + __popc(m & ((1u << lane) - 1u));
      mout[pos] = make_int4(li, s_match[li], 0, __float_as_int((float)s_odist[li]));
    }
    if (li < nL) {
      double x = 0.0, y = 0.0, z = 0.0;
      if (has) {
        const float* kl = kp_ptr(F, frame, 0, li);
        z = (double)g_depth[(long long)frame * cap + li];
        x = __ddiv_rn(__dmul_rn(__dsub_rn((double)kl[0], S.cx), z), S.fx);
        y = __ddiv_rn(__dmul_rn(__dsub_rn((double)kl[1], S.cy), z), S.fy);
      }
      pout[3 * li] = x;
      pout[3 * li + 1] = y;
      pout[3 * li + 2] = z;
      vout[li] = has ? 1 : 0;
    }
    __syncthreads();
    if (tid == 0) {
      int tot = 0;
      for (int w2 = 0; w2 < 32; ++w2) tot += s_part[w2];
      s_misc[1] += tot;
    }
    __syncthreads();
  }
  if (tid == 0) hdr->n_matches = s_misc[1];
}

}  // namespace vk
