// Internal parameter blocks shared by the host pipeline (vslam_b200.cu) and the kernels.
// Passed BY VALUE as __grid_constant__ kernel parameters (constant bank, no extra load latency).
#pragma once
#include <stdint.h>

#define VSLAM_MAX_LEVELS 12
#ifndef VSLAM_FAST_ROWS
#define VSLAM_FAST_ROWS 26      // emission rows per FAST strip (+ 2 dilation rows = 4 phase-A segments of 7)
#endif
#define VSLAM_RESIZE_ROWS 32    // destination rows of one resize lane item (4 columns wide)
#define VSLAM_BLUR_ROWS 96      // max output rows of one blur lane item (8 columns wide); 64 / 96 / 128 measured: 95.1k / 95.7k / 96.3k frames/s, 128 costs end-to-end throughput
#define VSLAM_DESC_BYTES 32

// 8-byte corner record permuted by the selection replay: hi = float bits of the response, lo = (y<<16)|x.
typedef unsigned long long vrec_t;

struct LevelP {
  int w, h, pitch;        // pitch: bytes per row, multiple of 16
  int pyr_off;            // byte offset inside the per-image pyramid slab (levels >= 1; level 0 = the input image)
  int blur_off;           // byte offset inside the per-image blurred slab (all levels)
  float scale;            // layerScale[l]
  int nfeat;              // nfeaturesPerLevel[l]
  int xcoef_off, ycoef_off;  // u32 offsets into the resize coefficient table (levels >= 1)
  int ex0, ex1, ey0, ey1; // emission region of FAST corners after runByImageBorder: [ex0,ex1) x [ey0,ey1)
  int fast_strip0, fast_nstrips;   // strips of VSLAM_FAST_ROWS emission rows
  int strip_cap;          // corner records per strip (worst case after NMS)
  int rec_off;            // record offset of this level inside the per-image strip-record slab
  int blur_item0, blur_ncg, blur_nstr, blur_rs;  // blur lane items: first item (multiple of 32), 8-column groups,
                                                 // row strips, rows per strip
  int pyr_cg0, pyr_ncg, pyr_y0, pyr_y1;          // 4-column groups and rows [y0, y1) of this level that k_resize computes: what
                                                 // FAST, Harris, IC angle, the (restricted) blur and the next level read
  int blur_cg0, blur_y0, blur_y1;                // first column group and row range [y0, y1) that are blurred: only what
                                                 // BRIEF can sample (>= edge_threshold - 18 px from every border)
};

struct PyrP {
  int nlevels;
  int eyes;               // images per frame in this launch (2 = stereo, 1 = mono)
  int in_pitch;           // row pitch of level-0 images
  int fast_threshold;
  int total_fast_strips;
  int total_blur_items;
  int kp_cap;             // per-image keypoint capacity
  int sel_smem_cap;       // records the selection kernel can hold in shared memory
  long long in_frame_stride;   // bytes between consecutive frames in the input (per eye base)
  long long pyr_img_stride;    // bytes per image in the pyramid slab
  long long blur_img_stride;   // bytes per image in the blurred slab
  long long rec_img_stride;    // records per image in the strip-record slab
  LevelP lv[VSLAM_MAX_LEVELS];
};

// Per-frame result header.
struct ResHdr {
  int n_kp[2];
  int n_matches;
  int status;
  int pad[4];
};

// Results of a chunk, structure-of-arrays over frames: array X of frame f starts at X + f * X_stride bytes.
// The strides equal the natural [frames][capacity] array strides, so a chunk can be copied with one D2H per array
// straight into caller arrays of that shape.
struct FrameLayout {
  ResHdr* hdr;                 // [frames]
  unsigned char* kp[2];        // cv::KeyPoint records, 28 B each
  unsigned char* desc[2];      // 32 B each
  unsigned char* matches;      // cv::DMatch records, 16 B each
  unsigned char* pts;          // 3 x f64 per LEFT keypoint
  unsigned char* valid;        // 1 B per LEFT keypoint
  long long kp_stride, desc_stride, matches_stride, pts_stride, valid_stride;
  int cap;                     // per-image keypoint capacity
};

struct StereoP {
  float max_disparity, min_disparity;  // (fx*b/0.1) as f32, (fx*b/40) as f32   stereo.rs:89-90
  double fxb, fx, fy, cx, cy;          // fx*baseline (f64 product), intrinsics
};
